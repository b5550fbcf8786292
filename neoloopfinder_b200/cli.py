"""``neoloop-caller`` on B200: same flags, same outputs (scripts/neoloop-caller of the reference).

What changes underneath:
* every assembly of a resolution is normalised and scored through the CUDA library, all assemblies
  of one GPU in ONE batch (``call_assemblies``); ``pipeline()`` keeps the reference's
  one-assembly signature (scripts/neoloop-caller:185-218) for drop-in use;
* the work is sharded over the visible GPUs by assembly (greedy longest-processing-time on the
  estimated pixel count); ``--nproc`` keeps its meaning for the CPU-side prologue;
* the model lookup bug of the reference (scripts/neoloop-caller:245 indexes the platform-keyed
  table with a resolution) is fixed to ``links[platform]``; without network the model file must
  already sit in ``--cachefolder`` under the reference's cache name
  ``peakachu_v2.res{R}.depth{D}.pkl`` (or be given with ``--model``);
* contact maps are opened with ``coolshim.load_cooler`` (.json synthetic spec, .npz arrays, or a
  real cooler URI when the ``cooler`` package exists).
"""
from __future__ import annotations

import argparse
import logging
import logging.handlers
import os
import sys
import traceback
from typing import Dict, List, Optional

import numpy as np

__version__ = "0.5.0+b200.1"

log = logging.getLogger(__name__)


# ----------------------------------------------------------------------------------------------
# scoring of assemblies
# ----------------------------------------------------------------------------------------------
from .sharding import assembly_weight, estimate_pixels, lpt_shards  # noqa: E402,F401 (re-exported)


def _regress_chunk(tasks, expected):
    from .callers import linear_regression_fit
    return [linear_regression_fit(le, expected, 5) for le in tasks]


ASSEMBLY_BUDGET_MB = 16384      # device memory one sub-batch of assemblies may hold (NLF_ASSEMBLY_BUDGET_MB)
MAX_JOBS_PER_BATCH = 4096       # below the library's NLF_MAX_JOBS (16384)


def _assembly_bytes(n_bins):
    """Device footprint of one assembly until its sub-batch is closed: the n x n fusion matrix, its band
    (416 doubles per row) and the probability plane (13 * 32 doubles per row)."""
    return 8 * n_bins * n_bins + 8 * n_bins * (416 + 416)


def _score_chunk(clr, live, assemblies, rsize, balance, prob, mmp, nopool, forest, expected, ctx, stats, nproc, out,
                 collect=None):
    """Normalise, score and annotate one sub-batch of valid assemblies [(ID, complexSV)]."""
    from .assembly import complexSV
    from .scoring import ScoreBatch, score_and_pool
    res = clr.binsize
    batch = None
    try:
        for _ID, fu in live:
            fu._build_matrix()
        # normalisation: diagonal means on the GPU for every assembly, then all regressions, then rescale
        task_lists = [fu.regression_tasks() for _ID, fu in live]
        flat = [t for tl in task_lists for t in tl]
        if nproc > 1 and len(flat) > 8:
            from functools import partial
            from .util import run_parallel
            chunk = max(1, len(flat) // (nproc * 4))
            chunks = [flat[k:k + chunk] for k in range(0, len(flat), chunk)]
            parts = run_parallel(nproc, partial(_regress_chunk, expected=expected), chunks)
            fitted = [r for part in parts for r in part]
        else:
            fitted = _regress_chunk(flat, expected)
        k = 0
        for (_ID, fu), tl in zip(live, task_lists):
            fu.apply_regressions(fitted[k:k + len(tl)])
            k += len(tl)
            fu.remove_gaps()
        exp_arr = np.r_[[expected[i] for i in sorted(expected)]]
        batch = ScoreBatch(4, 400, ctx)               # upper = 400*res // res (scripts/neoloop-caller:202)
        chains = []
        for _ID, fu in live:
            asm = fu._device_assembly()
            batch.add_assembly(asm)
            chain_of = np.array([fu.chains[k] for k in range(fu._n_bins)])
            chains.append(chain_of[asm.kept_index()])     # chains[index_map[i]] per gap-free bin
            if collect is not None:
                collect[_ID] = (asm.gap_free(), chains[-1])
        loops_by_job, (n_cand, n_scored, n_don) = score_and_pool(batch, forest, exp_arr, prob, res, mmp, chains, nopool)
        if stats is not None:
            stats["candidates"] = stats.get("candidates", 0) + int(n_cand.sum())
            stats["scored"] = stats.get("scored", 0) + int(n_scored.sum())
            stats["donuts"] = stats.get("donuts", 0) + int(n_don.sum())
            stats["sub_batches"] = stats.get("sub_batches", 0) + 1

        for job, (ID, fu) in enumerate(live):
            li, lj, lp = loops_by_job[job]
            loop_index = [(int(a), int(b), float(p)) for a, b, p in zip(li, lj, lp)]
            loops = fu.index_to_coordinates(loop_index)
            # keep calls inside the ORIGINAL (flexible) assembly only; geometry without re-fetching
            fu2 = complexSV(clr, assemblies[ID], span=rsize, col=balance, protocol='insitu', expected_values=expected)
            fu2.reorganize_matrix(map_only=True)
            out[ID] = {k: (ID, v[0], v[1]) for k, v in loops.items()
                       if (k[0], k[1]) in fu2.Map and (k[3], k[4]) in fu2.Map}
    finally:
        # the sub-batch's device memory goes back before the next one is built
        for _ID, fu in live:
            if fu._asm is not None:
                fu._asm.close()
                fu._asm = None
        if batch is not None:
            batch.close()


def call_assemblies(clr, ids, assemblies, rsize, balance, prob, mmp, nopool, model, expected, ctx=None,
                    stats=None, nproc=1, budget_mb=None, collect=None):
    """pipeline() for many assemblies on one GPU, batched.  Returns {ID: final_dict or None}.
    ``collect`` (a dict) receives {ID: (gap-free matrix on the host, block id per gap-free bin)}.

    Assemblies are validated from their geometry alone (an invalid one costs no device memory) and
    processed in sub-batches bounded by a device-memory budget and by the library's job limit; every
    sub-batch is closed before the next one is built, so an assembly file of any length runs in
    bounded memory.  Results are per assembly: the split does not change them.

    The per-block-pair regressions (HuberRegressor, host-bound, ~5 ms each, ~18 per simple
    assembly) dominate the wall time once the scoring runs on the GPU; with nproc > 1 they are
    farmed out to loky worker processes, as the reference does for whole assemblies."""
    from . import _lib
    from .assembly import complexSV
    from .forest import DeviceForest, load_forest

    ctx = ctx or _lib.Context.get()
    forest = model if isinstance(model, DeviceForest) else load_forest(model, ctx)
    if budget_mb is None:
        budget_mb = int(os.environ.get("NLF_ASSEMBLY_BUDGET_MB", ASSEMBLY_BUDGET_MB))
    budget = max(1, int(budget_mb)) << 20
    out = {}
    live, held = [], 0
    for ID in ids:
        fu = complexSV(clr, assemblies[ID], span=rsize, col=balance, protocol='insitu',
                       expected_values=expected, flexible=False)
        fu._ctx = ctx
        fu.reorganize_matrix(map_only=True)
        if len(fu.Map) != fu._n_bins:
            out[ID] = None                         # invalid assembly (scripts/neoloop-caller:196-197)
            continue
        need = _assembly_bytes(fu._n_bins)
        if live and (held + need > budget or len(live) >= MAX_JOBS_PER_BATCH):
            _score_chunk(clr, live, assemblies, rsize, balance, prob, mmp, nopool, forest, expected, ctx, stats, nproc, out, collect)
            live, held = [], 0
        live.append((ID, fu))
        held += need
    if live:
        _score_chunk(clr, live, assemblies, rsize, balance, prob, mmp, nopool, forest, expected, ctx, stats, nproc, out, collect)
    return out


def pipeline(clr, index, assemblies, rsize, balance, prob, mmp, nopool, models_by_res, expected):
    """One assembly, the reference's signature (scripts/neoloop-caller:185-218)."""
    res = clr.binsize
    return call_assemblies(clr, [index], assemblies, rsize, balance, prob, mmp, nopool, models_by_res[res],
                           expected)[index]


def _device_worker(dev, clr, ids, assemblies, rsize, balance, prob, mmp, nopool, model_path, expected, nproc, budget_mb,
                   child=False):
    """Body of one per-GPU worker process (or thread): its own CUDA context, forest copy and pixel table.
    ``child``: running in a spawned worker process, which must stop its own loky pool before it returns
    (``util.shutdown_workers``)."""
    from . import _lib
    try:
        ctx = _lib.Context.get(dev)
        st = {}
        r = call_assemblies(clr, ids, assemblies, rsize, balance, prob, mmp, nopool, model_path, expected, ctx, st, nproc,
                            budget_mb)
        st["launches"] = ctx.launch_count()
        return r, st
    finally:
        if child:
            from .util import shutdown_workers
            shutdown_workers()


def call_resolution(clr, assemblies: Dict[str, str], rsize, balance, prob, mmp, nopool, model_path, expected,
                    devices: Optional[List[int]] = None, stats=None, nproc=1, workers: Optional[str] = None,
                    budget_mb=None):
    """All assemblies of one resolution, sharded over GPUs by assembly (no collective: the
    per-assembly results are KB-sized dictionaries merged on the host).

    With more than one device every GPU is driven by its own worker PROCESS (spawned, so each has its
    own interpreter and CUDA context: the Python-side geometry, dictionary building and regression
    glue of different GPUs never share a GIL) -- the reference's own fan-out is one loky process per
    assembly (scripts/neoloop-caller:155-159).  ``workers='thread'`` (or NLF_WORKERS=thread) keeps
    everything in this process, one thread per GPU; it is also what a non-picklable model handle needs."""
    from . import _lib
    from .forest import DeviceForest
    if devices is None:
        devices = list(range(_lib.device_count()))
    devices = list(dict.fromkeys(int(d) for d in devices))          # a context must not be driven by two workers
    if not devices:
        raise RuntimeError("no CUDA device visible: neoloopfinder_b200 has no CPU fallback")
    ids = list(assemblies)
    res = clr.binsize
    # cost estimate from the assembly geometry only
    weights = [assembly_weight(assemblies[ID], rsize, res, clr) for ID in ids]
    shards = lpt_shards(weights, len(devices))
    jobs = [(d, [ids[k] for k in s]) for d, s in zip(devices, shards) if s]
    per_dev_nproc = max(1, nproc // max(1, len(jobs)))
    mode = workers or os.environ.get("NLF_WORKERS", "process")
    if isinstance(model_path, DeviceForest):
        mode = "thread"
    results: Dict[str, Optional[dict]] = {}
    per_dev = []

    def run(dev, mine):
        return _device_worker(dev, clr, mine, assemblies, rsize, balance, prob, mmp, nopool, model_path, expected,
                              per_dev_nproc, budget_mb)

    if len(jobs) <= 1:
        per_dev = [run(*j) for j in jobs]
    elif mode == "thread":
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            per_dev = list(ex.map(lambda j: run(*j), jobs))
    else:
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor
        with ProcessPoolExecutor(max_workers=len(jobs), mp_context=mp.get_context("spawn")) as ex:
            futs = [ex.submit(_device_worker, d, clr, mine, assemblies, rsize, balance, prob, mmp, nopool, model_path,
                              expected, per_dev_nproc, budget_mb, True) for d, mine in jobs]
            per_dev = [f.result() for f in futs]
    for r, st in per_dev:
        results.update(r)
        if stats is not None:
            for k, v in st.items():
                stats[k] = stats.get(k, 0) + v
    return [results[ID] for ID in ids]


def merge_results(results):
    """scripts/neoloop-caller:161-170."""
    cache = {}
    for tmp in results:
        if tmp is None:
            continue
        for key in tmp:
            cache.setdefault(key, []).append(tmp[key])
    return cache


# ----------------------------------------------------------------------------------------------
# model lookup (scripts/neoloop-caller:220-270, fixed and offline-aware)
# ----------------------------------------------------------------------------------------------
def match_digital_values(arr, v):
    arr = list(arr)
    return arr[int(np.argmin(np.abs(v - np.r_[arr])))]


def locate_peakachu_models(cachefolder, clrs_by_res, platform='Hi-C'):
    """{res: model path}.  Picks the model whose resolution and intra-chromosomal depth are
    nearest, as the reference does (with the platform lookup fixed), then expects the file in the
    cache folder under the reference's cache name (no network in this build)."""
    from .model_table import model_table
    table = model_table(platform)
    out = {}
    for res, clr in clrs_by_res.items():
        m_res = match_digital_values(table, res)
        info = clr.info
        if 'sum' in info:
            depth = info['sum']
        elif info['nnz'] < 100000000:
            depth = clr.pixels()[:]['count'].values.sum()
        else:
            depth = info['nnz'] * 2.14447683
        depth = 3031042417 / clr.chromsizes.sum() * depth
        depth = res / m_res * depth
        intra = depth * 0.73256754
        m_depth = match_digital_values(table[m_res], intra)
        path = os.path.join(cachefolder, table[m_res][m_depth])
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found. This build has no network access: place the Peakachu model "
                f"({platform}, {m_res} bp, depth {m_depth}) there or pass --model.")
        out[res] = path
    return out


# ----------------------------------------------------------------------------------------------
# command line
# ----------------------------------------------------------------------------------------------
def getargs(argv=None):
    parser = argparse.ArgumentParser(description='''Identify novel loop interactions across SV
                                     points.''', formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('-v', '--version', action='version', version=' '.join(['%(prog)s', __version__]),
                        help='Print version number and exit.')
    parser.add_argument('-O', '--output', help='Path to the output file.')
    parser.add_argument('-H', '--hic', nargs='+', help='''List of cooler URIs. If URIs at multiple
                        resolutions are provided, the program will first detect neo-loops from
                        each individual resolution, and then combine results from all resolutions
                        in a non-redundant way. If an interaction is detected as a loop in multiple
                        resolutions, only the one with the highest resolution will be recorded.''')
    parser.add_argument('--assembly', help='''The assembled SV list outputed by assemble-complexSVs.''')
    parser.add_argument('--cachefolder', default='.cache', help='''Path to a folder to place the pre-trained
                        Peakachu models.''')
    parser.add_argument('-R', '--region-size', default=3000000, type=int,
                        help='''The extended genomic span of SV break points.(bp)''')
    parser.add_argument('--balance-type', default='CNV', choices=['CNV', 'ICE', 'RAW'], help='Normalization method.')
    parser.add_argument('--platform', default='Hi-C',
                        choices=['Hi-C', 'CTCF-ChIAPET', 'Pol2-ChIAPET', 'H3K27ac-HiChIP', 'H3K4me3-HiChIP',
                                 'CTCF-HiChIP', 'SMC1A-HiChIP', 'HiCAR', 'TrAC-loop'],
                        help='''Name of the experimental platform you used to generate your chromatin contact data.''')
    parser.add_argument('--prob', type=float, default=0.9, help='Probability threshold.')
    parser.add_argument('--no-clustering', action='store_true', help='No pooling will be performed if specified.')
    parser.add_argument('--min-marginal-peaks', type=int, default=2,
                        help='''Minimum marginal number of loops when detecting loop anchors.''')
    parser.add_argument('--nproc', default=1, type=int, help='Number of worker processes.')
    parser.add_argument('--logFile', default='neoloop.log', help='''Logging file name.''')
    # additions of this build (all optional)
    parser.add_argument('--model', nargs='+', default=None,
                        help='''Model pickle(s), one per -H matrix in the same order; overrides the cache lookup.''')
    parser.add_argument('--gpus', default=None,
                        help='''Comma-separated CUDA device indices (default: every visible device).''')
    parser.add_argument('--genome-wide', action='store_true',
                        help='''Also call loops on every chromosome outside any SV assembly (each chromosome is scored as
                        a trivial, junction-free assembly). --assembly becomes optional.''')
    parser.add_argument('--chroms', nargs='+', default=None,
                        help='''Chromosomes of the genome-wide call (default: all except unplaced contigs, chrM and EBV).''')
    commands = list(sys.argv[1:] if argv is None else argv)
    if not commands:
        commands.append('-h')
    return parser.parse_args(commands), commands


def run(argv=None):
    args, commands = getargs(argv)
    if commands[0] in ['-h', '-v', '--help', '--version']:
        return
    logger = logging.getLogger()
    logger.setLevel(10)
    console = logging.StreamHandler()
    filehandler = logging.handlers.RotatingFileHandler(args.logFile, maxBytes=100000, backupCount=5)
    console.setLevel('INFO')
    filehandler.setLevel('INFO')
    formatter = logging.Formatter(fmt='%(name)-25s %(levelname)-7s @ %(asctime)s: %(message)s',
                                  datefmt='%m/%d/%y %H:%M:%S')
    console.setFormatter(formatter)
    filehandler.setFormatter(formatter)
    logger.addHandler(console)
    logger.addHandler(filehandler)
    arglist = ['# ARGUMENT LIST:',
               '# Output Path = {0}'.format(args.output),
               '# Cache Folder = {0}'.format(args.cachefolder),
               '# SV assembly = {0}'.format(args.assembly),
               '# Cooler List = {0}'.format(args.hic),
               '# Extended Genomic Span = {0}bp'.format(args.region_size),
               '# Balance Type = {0}'.format(args.balance_type),
               '# Probability threshold = {0}'.format(args.prob),
               '# No pooling = {0}'.format(args.no_clustering),
               '# Minimum marginal peaks = {0}'.format(args.min_marginal_peaks),
               '# Number of Processes = {0}'.format(args.nproc),
               '# Log file name = {0}'.format(args.logFile)]
    logger.info('\n' + '\n'.join(arglist))

    from .callers import combine_annotations
    from .coolshim import load_cooler
    from .util import autosomes, calculate_expected

    balance = {'CNV': 'sweight', 'ICE': 'weight', 'RAW': False}[args.balance_type]
    try:
        clrs = {}
        order = []
        for path in args.hic:
            lib = load_cooler(path)
            clrs[lib.binsize] = lib
            order.append(lib.binsize)
        cachefolder = os.path.abspath(os.path.expanduser(args.cachefolder))
        os.makedirs(cachefolder, exist_ok=True)
        if args.model:
            if len(args.model) != len(order):
                raise ValueError('--model needs one path per -H matrix')
            models_by_res = dict(zip(order, args.model))
        else:
            models_by_res = locate_peakachu_models(cachefolder, clrs, args.platform)
        devices = None if args.gpus is None else [int(t) for t in args.gpus.split(',')]

        logger.info('Load assembled SVs')
        lines = {}
        if args.assembly is None and not args.genome_wide:
            raise ValueError('--assembly is required unless --genome-wide is given')
        if args.assembly is not None:
            with open(args.assembly, 'r') as source:
                for line in source:
                    parse = line.rstrip().split()
                    if parse:
                        lines[parse[0]] = '\t'.join(parse[1:])

        byres = {}
        for res in sorted(clrs, reverse=True):
            logger.info('Current resolution: {0}'.format(res))
            clr = clrs[res]
            logger.info('Calculate the global average contact frequencies at each genomic distance ...')
            expected = calculate_expected(clr, autosomes(clr.chromnames), maxdis=5000000, balance=balance,
                                          nproc=args.nproc)
            logger.info('Done')
            stats = {}
            results = []
            if lines:
                results = call_resolution(clr, lines, args.region_size, balance, args.prob, args.min_marginal_peaks,
                                          args.no_clustering, models_by_res[res], expected, devices, stats, args.nproc)
            logger.info('Scored {0} pixels ({1} candidates, {2} above the threshold)'.format(
                stats.get('scored', 0), stats.get('candidates', 0), stats.get('donuts', 0)))
            logger.info('Merge loops across assemblies ...')
            byres[res] = merge_results(results)
            if args.genome_wide:
                from .genome import call_genome, genome_lines
                names = args.chroms or [c for c in clr.chromnames if not any(t in c for t in ('_', 'M', 'EBV'))]
                logger.info('Genome-wide call on {0} chromosomes ...'.format(len(names)))
                gstats = {}
                calls = call_genome(clr, models_by_res[res], expected, args.prob, args.min_marginal_peaks,
                                    args.no_clustering, names, balance, devices, gstats)
                logger.info('Scored {0} pixels on {1} bins'.format(gstats.get('scored', 0), gstats.get('bins', 0)))
                for key, labels in genome_lines(calls, res).items():
                    byres[res].setdefault(key, []).extend(labels)

        if len(clrs) > 1:
            logger.info('Combine loops from multiple resolutions')
        loop_list = combine_annotations(byres)
        logger.info('Output ...')
        with open(args.output, 'w') as out:
            for line in loop_list:
                out.write('\t'.join(line) + '\n')
        logging.info('Done')
    except:  # noqa: E722 - same contract as the reference: traceback to the log file, exit code 1
        traceback.print_exc(file=open(args.logFile, 'a'))
        sys.exit(1)


if __name__ == '__main__':
    run()
