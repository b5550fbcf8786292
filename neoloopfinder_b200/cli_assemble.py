"""``assemble-complexSVs`` on B200: the command line of the reference script
(scripts/assemble-complexSVs) over ``assembly.assembleSV``.

Same flags and output (``<prefix>.assemblies.txt``: ``A#`` allele chains, then ``C#`` single SVs).  The
reference script reads ``args.protocol`` although it defines no such flag and stops there
(scripts/assemble-complexSVs:93); here ``--protocol`` exists (default ``insitu``), everything else is as is.
"""
from __future__ import annotations

import argparse
import logging
import logging.handlers
import os
import sys
import traceback

from . import __version__

PLATFORMS = ['Hi-C', 'CTCF-ChIAPET', 'Pol2-ChIAPET', 'H3K27ac-HiChIP', 'H3K4me3-HiChIP', 'CTCF-HiChIP',
             'SMC1A-HiChIP', 'HiCAR', 'TrAC-loop']


def getargs(argv=None):
    parser = argparse.ArgumentParser(description='Assemble complex SVs given an individual SV list.',
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('-v', '--version', action='version', version=' '.join(['%(prog)s', __version__]),
                        help='Print version number and exit.')
    parser.add_argument('-O', '--output', help='Output prefix.')
    parser.add_argument('-B', '--break-points', help='Path to a TXT file containing pairs of break points.')
    parser.add_argument('-H', '--hic', nargs='+', help='''List of cooler URIs. With several resolutions, complex SVs
                        are detected at each of them and merged in a non-redundant way.''')
    parser.add_argument('-R', '--region-size', default=5000000, type=int,
                        help='The extended genomic span of SV break points.(bp)')
    parser.add_argument('--minimum-size', default=500000, type=int,
                        help='Intra-chromosomal SVs smaller than this size are ignored.')
    parser.add_argument('--balance-type', default='CNV', choices=['CNV', 'ICE', 'RAW'], help='Normalization method.')
    parser.add_argument('--platform', default='Hi-C', choices=PLATFORMS,
                        help='Name of the experimental platform that produced the contact data.')
    parser.add_argument('--protocol', default='insitu', help='Experimental protocol tag handed to the SV objects.')
    parser.add_argument('--nproc', default=1, type=int, help='Number of worker processes.')
    parser.add_argument('--logFile', default='assembleSVs.log', help='Logging file name.')
    commands = list(sys.argv[1:] if argv is None else argv)
    if not commands:
        commands.append('-h')
    return parser.parse_args(commands), commands


def parse_assemblies(path):
    """{tuple of SV tokens: [bound1, bound2]} for the A# and the C# lines of an assembly file."""
    chains, singles = {}, {}
    with open(path) as source:
        for line in source:
            f = line.rstrip().split()
            if f:
                (chains if f[0].startswith('A') else singles if f[0].startswith('C') else {})[tuple(f[1:-2])] = f[-2:]
    return chains, singles


def _contiguous_in(p1, p2):
    pos = []
    for item in p1:
        if item not in p2:
            return False
        pos.append(p2.index(item))
    return all(b - a == 1 for a, b in zip(pos, pos[1:]))


def combine(by_res):
    """Union over resolutions, the finest one first; a chain contained in a longer kept chain is dropped
    (scripts/assemble-complexSVs:206-233)."""
    finest = min(by_res)
    chains, singles = parse_assemblies(by_res[finest])
    for r in by_res:
        if r != finest:
            more_chains, more_singles = parse_assemblies(by_res[r])
            for k, v in more_chains.items():
                chains.setdefault(k, v)
            for k, v in more_singles.items():
                singles.setdefault(k, v)
    kept = []
    for _n, p in sorted(((len(k), k) for k in chains), reverse=True):
        if not any(_contiguous_in(p, v) for v in kept):
            kept.append(p)
    return {k: chains[k] for k in kept}, singles


def run(argv=None):
    args, commands = getargs(argv)
    if commands[0] in ['-h', '-v', '--help', '--version']:
        return
    logger = logging.getLogger()
    logger.setLevel(10)
    console = logging.StreamHandler()
    filehandler = logging.handlers.RotatingFileHandler(args.logFile, maxBytes=100000, backupCount=5)
    fmt = logging.Formatter(fmt='%(name)-25s %(levelname)-7s @ %(asctime)s: %(message)s', datefmt='%m/%d/%y %H:%M:%S')
    for h in (console, filehandler):
        h.setLevel('INFO')
        h.setFormatter(fmt)
        logger.addHandler(h)
    logger.info('\n' + '\n'.join([
        '# ARGUMENT LIST:', '# Output Prefix = {0}'.format(args.output), '# Break Points = {0}'.format(args.break_points),
        '# Minimum fragment size = {0}bp'.format(args.minimum_size), '# Cooler URI = {0}'.format(args.hic),
        '# Extended Genomic Span = {0}bp'.format(args.region_size), '# Balance Type = {0}'.format(args.balance_type),
        '# Experimental protocol = {0}'.format(args.protocol), '# Number of Processes = {0}'.format(args.nproc),
        '# Log file name = {0}'.format(args.logFile)]))

    from .assembly import assembleSV
    from .coolshim import load_cooler
    from .util import autosomes, calculate_expected

    balance = {'CNV': 'sweight', 'ICE': 'weight', 'RAW': False}[args.balance_type]
    try:
        clrs = {}
        for path in args.hic:
            lib = load_cooler(path)
            clrs[lib.binsize] = lib
        by_res = {}
        for res, clr in clrs.items():
            logger.info('Current resolution: {0}'.format(res))
            logger.info('Calculate the global average contact frequencies at each genomic distance ...')
            expected = calculate_expected(clr, autosomes(clr.chromnames), maxdis=5000000, balance=balance, nproc=args.nproc)
            logger.info('Filtering SVs by checking distance decay of chromatin contacts across SV breakpoints ...')
            work = assembleSV(clr, args.break_points, span=args.region_size, col=balance, n_jobs=args.nproc,
                              protocol=args.protocol, minIntra=args.minimum_size, expected_values=expected)
            logger.info('{0} SVs left'.format(len(work.queue)))
            logger.info('Building SV connecting graph ...')
            work.build_graph()
            logger.info('Discovering and re-ordering complex SVs ...')
            work.find_complexSV()
            logger.info('Called {0} connected assemblies'.format(len(work.alleles)))
            by_res[res] = '{0}.assemblies.{1}.txt'.format(args.output, res)
            work.output_all(by_res[res])

        outfil = '{0}.assemblies.txt'.format(args.output)
        if len(by_res) == 1:
            os.replace(next(iter(by_res.values())), outfil)
        else:
            logger.info('Merge complex SV results from multiple resolutions ...')
            chains, singles = combine(by_res)
            with open(outfil, 'w') as out:
                ranked = sorted(((len(k),) + k + tuple(chains[k]) for k in chains), reverse=True)
                for i, rec in enumerate(ranked):
                    out.write('\t'.join(['A{0}'.format(i)] + list(rec[1:])) + '\n')
                for i, k in enumerate(singles):
                    out.write('\t'.join(['C{0}'.format(i)] + list(k) + list(singles[k])) + '\n')
            for path in by_res.values():
                os.remove(path)
        logger.info('Done')
    except:  # noqa: E722 - same contract as the reference: traceback to the log file, exit code 1
        traceback.print_exc(file=open(args.logFile, 'a'))
        sys.exit(1)


if __name__ == '__main__':
    run()
