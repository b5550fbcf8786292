"""Drop-in for ``neoloop.assembly.complexSV`` on the neoloop-caller path (neoloop/assembly.py:278-596).

Host side: parsing of the assembly line, block geometry, the cooler fetches that build the
rearranged matrix, coordinate bookkeeping and the tiny per-block-pair regression.  Device side
(libnlf_b200): every pass over the n x n matrix -- masked per-diagonal means in numpy's summation
order, the per-block rescale, the gap-column sums and the gap-free compaction.

``assembleSV`` / ``filterSV`` / ``filterAssembly`` (the inner loop of
``assemble-complexSVs``, neoloop/assembly.py:12-275) reuse the same kernels: every SV and every candidate
chain of SVs is tested by the distance decay of the contacts it induces.
"""
from __future__ import annotations

import logging

import numpy as np

from .callers import Fusion
from .scoring import DeviceAssembly
from .util import find_chrom_pre

log = logging.getLogger(__name__)

__all__ = ["complexSV", "assembleSV", "filterSV", "filterAssembly"]

from .callers import _MAXDIS_LADDER                       # assembly.py:513


class complexSV(Fusion):

    def __init__(self, clr, candidate, span=5000000, col='sweight', protocol='insitu',
                 flexible=True, slopes={}, expected_values=None):
        self.clr = clr
        self.res = clr.binsize
        self.protocol = protocol
        self.pre = find_chrom_pre(list(clr.chromnames))
        self.span = span
        self.flexible = flexible
        self.slopes = slopes
        self.balance_type = col if col in ('weight', 'sweight') else False
        self.chroms = {c: [0, clr.chromsizes[c]] for c in clr.chromnames}
        self._asm = None
        self._gap_free = None
        self._hcm = None
        self._scale_plan = None
        self._fusion_matrix = None

        self.parse_input(candidate)
        self.tb, self.to = [], []
        last = len(self.sv_list) - 1
        for k, sv in enumerate(self.sv_list):
            # only the outermost ends honour the user-given bounds (assembly.py:303-318)
            left = self.bounds[0] if k == 0 else None
            right = self.bounds[1] if k == last else None
            intervals, directs = self.get_single_block(sv, left_bound=left, right_bound=right)
            self.tb.extend(intervals)
            self.to.extend(directs)

        if expected_values is None:
            self.load_expected()
        else:
            self.expected = expected_values

    # ------------------------------------------------------------------ parsing / geometry
    def parse_input(self, candidate):
        """'type,c1,p1,s1,c2,p2,s2 ... chrom,bound1 chrom,bound2' (assembly.py:323-335)."""
        fields = candidate.split()
        self.sv_list = []
        for tok in fields[:-2]:
            _, c1, p1, s1, c2, p2, s2 = tok.split(',')
            self.sv_list.append((c1, c2, int(p1), int(p2), s1, s2))
        b1 = fields[-2].split(',')
        b2 = fields[-1].split(',')
        self.bounds = [(b1[0], int(b1[1])), (b2[0], int(b2[1]))]

    def _piece(self, chrom, pos, keep_left, bound):
        """Interval retained next to a breakpoint: the side left of ``pos`` when ``keep_left``,
        out to ``bound`` (flexible assemblies) or to +-span clipped to the chromosome."""
        lo, hi = self.chroms[chrom]
        if self.flexible and bound is not None:
            return [chrom, bound[1], pos] if keep_left else [chrom, pos, bound[1]]
        if keep_left:
            return [chrom, max(lo, pos - self.span), pos]
        return [chrom, pos, min(pos + self.span, hi)]

    def get_single_block(self, sv, left_bound=None, right_bound=None):
        """Two intervals and their traversal directions for one SV (assembly.py:337-385):
        '+' at a breakpoint keeps the side left of it, '-' the side right of it; the second
        piece is traversed 5'->3' ('+') exactly when it lies right of its breakpoint."""
        c1, c2, p1, p2, s1, s2 = sv
        c1, c2 = self.pre + c1, self.pre + c2
        directions = [s1, '-' if s2 == '+' else '+']
        intervals = [self._piece(c1, p1, s1 == '+', left_bound),
                     self._piece(c2, p2, s2 != '-', right_bound)]
        return intervals, directions

    def get_bias(self, r):
        col = 'weight' if self.balance_type == False else self.balance_type   # noqa: E712 (reference semantics)
        bias = self.clr.bins().fetch(r)[col].values
        bias = np.array(bias, dtype=np.float64, copy=True)
        bias[np.isnan(bias)] = 0
        return bias

    def get_matrix(self, r1, r2):
        M = self.clr.matrix(balance=self.balance_type).fetch(r1, r2)
        if self.balance_type:
            M[np.isnan(M)] = 0
        return M

    def reorganize_matrix(self, map_only=False):
        """Blocks, orientations, bin index, coordinate maps and the rearranged upper-triangular
        matrix (assembly.py:408-492).  I/O bound: (k+1)(k+2)/2 region fetches.

        ``map_only=True`` builds the geometry (Map, reverse, bounds, index, chains) without
        fetching any contact data: what the reference's second, flexible ``complexSV`` is used
        for at scripts/neoloop-caller:210-216."""
        tb, to = self.tb, self.to
        blocks, orients = [tb[0]], [to[0]]
        for i in range(1, len(tb) - 1, 2):
            # the piece leaving SV k and the piece entering SV k+1 form one middle block
            assert to[i] == to[i + 1]
            assert tb[i][0] == tb[i + 1][0]
            if to[i] == '+':
                a, b = tb[i][1], tb[i + 1][2]
                assert a < b
                blocks.append([tb[i][0], a, b])
            else:
                a, b = tb[i][2], tb[i + 1][1]
                assert a > b
                blocks.append([tb[i][0], b, a])
            orients.append(to[i])
        blocks.append(tb[-1])
        orients.append(to[-1])

        res = self.res
        index, Map, bounds = [], [], []
        start = 0
        for r, o in zip(blocks, orients):
            nbin = self.clr.bins().fetch(tuple(r))['chrom'].size
            index.append((start, start + nbin))
            first = r[1] // res * res
            coords = [(r[0], first + k * res) for k in range(nbin)]
            ids = range(start, start + nbin)
            lo_end, hi_end = (r[0], first), (r[0], r[2] // res * res)
            if o == '+':
                Map.extend(zip(coords, ids))
                bounds += [(start, lo_end), (start + nbin - 1, hi_end)]
            else:
                Map.extend(zip(coords[::-1], ids))
                bounds += [(start, hi_end), (start + nbin - 1, lo_end)]
            start += nbin
        forward = dict(Map)

        n = index[-1][1]
        self.Map = forward
        self.reverse = {forward[k]: k for k in forward}
        self.bounds = bounds
        self.orients = orients
        self.index = index
        self.blocks = blocks
        self.chains = {}
        for k, (lo, hi) in enumerate(index):
            for ri in range(lo, hi):
                self.chains[ri] = k
        self._asm = None
        self._gap_free = None
        self._hcm = None
        self._n_bins = n
        self._fusion_matrix = None
        if map_only:
            return
        self._build_matrix()

    def _build_matrix(self):
        """Second half of ``reorganize_matrix``: the contact data of the blocks laid out by the geometry
        half (assembly.py:458-492).  Separate so that a caller can validate the geometry
        (``len(Map) != n``: overlapping blocks) before any matrix is fetched or allocated."""
        blocks, orients, index, n = self.blocks, self.orients, self.index, self._n_bins
        if hasattr(self.clr, "device_pixels") and n > 0:
            # the sample's pixel table is resident in HBM; the rearranged matrix
            # and the bias vector are built there (no region fetches, no host n x n matrix, no upload)
            col = 'weight' if self.balance_type == False else self.balance_type   # noqa: E712
            px = self.clr.device_pixels(getattr(self, '_ctx', None), col)
            lo_hi = [self.clr.global_bins(*self.clr._bin_extent(tuple(r))) for r in blocks]
            self._asm = DeviceAssembly.from_pixels(px, [a for a, _b in lo_hi], [b for _a, b in lo_hi],
                                                   [o == '-' for o in orients], balance=bool(self.balance_type))
            self.bias = self._asm.bias()
            return
        M = np.zeros((n, n))
        bias_parts = []
        for j1, (r1, o1, i1) in enumerate(zip(blocks, orients, index)):
            b = self.get_bias(r1)
            bias_parts.append(b if o1 == '+' else b[::-1])
            for j2, (r2, o2, i2) in enumerate(zip(blocks, orients, index)):
                if j1 > j2:
                    continue
                sub = self.get_matrix(r1, r2)
                if o1 == '-':
                    sub = sub[::-1, :]
                if o2 == '-':
                    sub = sub[:, ::-1]
                M[i1[0]:i1[1], i2[0]:i2[1]] = sub

        self.fusion_matrix = np.triu(M)
        self.bias = np.concatenate(bias_parts) if bias_parts else np.r_[[]]

    @property
    def fusion_matrix(self):
        """np.triu of the rearranged matrix (assembly.py:483); fetched from the device on first use
        when the assembly was built from a resident pixel table."""
        if self._fusion_matrix is None and self._asm is not None:
            self._fusion_matrix = self._asm.fusion_matrix()
        return self._fusion_matrix

    @fusion_matrix.setter
    def fusion_matrix(self, M):
        self._fusion_matrix = M

    # ------------------------------------------------------------------ device plumbing
    def _device_assembly(self) -> DeviceAssembly:
        if self._asm is None:
            self._asm = DeviceAssembly(self.fusion_matrix, self.bias, self.index, getattr(self, '_ctx', None))
            if self._scale_plan is not None:
                self._asm.set_scales(*self._scale_plan)
        return self._asm

    # ------------------------------------------------------------------ normalisation
    def correct_heterozygosity(self):
        """Per-block-pair slopes against the global expected and the rescaled matrix
        (assembly.py:495-571).  GPU: masked diagonal means (once, to 2 Mb: the reference's six
        maxdis values are prefixes of it), rescale, symmetrise/triu.  Host: the regressions."""
        tasks = self.regression_tasks()
        self.apply_regressions([self.linear_regression(le, self.expected, min_samples=5) for le in tasks])

    def regression_tasks(self):
        """First half: diagonal means on the GPU, then the list of DISTINCT regression inputs
        (one local-expected dictionary per block pair and distinct maxdis prefix).  The fits are
        independent, tiny and host-bound (HuberRegressor / L-BFGS), so a caller may farm them out to
        worker processes and hand the results to ``apply_regressions`` in the same order."""
        asm = self._device_assembly()
        nb = len(self.blocks)
        res = self.res
        maxdiag = _MAXDIS_LADDER[-1] // res
        mean, _cnt = asm.diag_means(maxdiag, mink=3)
        self._reg_plan = []          # per pair: (j1, j2, [task index per ladder step], last local_exp)
        tasks = []
        p = 0
        for j1 in range(nb):
            for j2 in range(j1, nb):
                row = mean[p]
                p += 1
                kept = np.where(~np.isnan(row))[0]
                kept_keys = kept.tolist()
                kept_vals = list(row[kept])                  # numpy float64 scalars, as row[i] gives them
                memo, steps = {}, []
                local_exp = {}
                for maxdis in _MAXDIS_LADDER:
                    key = int(np.searchsorted(kept, maxdis // res, side="right"))     # prefixes: same length <=> same dictionary
                    if key not in memo:
                        memo[key] = len(tasks)
                        tasks.append(dict(zip(kept_keys[:key], kept_vals[:key])))
                    steps.append(memo[key])
                    local_exp = tasks[memo[key]]
                self._reg_plan.append((j1, j2, steps, local_exp))
        return tasks

    def apply_regressions(self, results):
        """Second half: pick the best fit per pair, rescale on the GPU (assembly.py:520-571)."""
        asm = self._device_assembly()
        nb = len(self.blocks)
        slopes, exps, pair_binary = {}, {}, {}
        for j1, j2, steps, local_exp in self._reg_plan:
            rs = np.r_[[results[k][0] for k in steps]]
            ss = [results[k][1] for k in steps]
            ws = [results[k][2] for k in steps]
            r = rs.max()
            s = ss[rs.argmax()]
            if any(ws):
                pair_binary[(j1, j2)] = 1 if (r > 0.6 and s > 0) else 0
            else:
                pair_binary[(j1, j2)] = 1
            slopes[(j1, j2)] = [r, s]
            exps[(j1, j2)] = local_exp

        op = np.zeros((nb, nb), dtype=np.int32)
        val = np.ones((nb, nb), dtype=np.float64)
        for j1 in range(nb):
            for j2 in range(j1, nb):
                ms = max(slopes[(j1, j1)][1], slopes[(j2, j2)][1])
                if (j1, j2) in self.slopes:
                    op[j1, j2], val[j1, j2] = 1, self.slopes[(j1, j2)]
                elif slopes[(j1, j2)][0] > 0.6 and slopes[(j1, j2)][1] > 0 and ms > 0:
                    if j1 == j2:
                        op[j1, j2], val[j1, j2] = 1, slopes[(j1, j2)][1]
                    else:
                        op[j1, j2], val[j1, j2] = 2, min(1 / slopes[(j1, j2)][1], 5 / ms)
        self._scale_plan = (op, val)
        asm.set_scales(op, val)
        self._hcm = None
        self._gap_free = None

        self.slopes_ = slopes
        self.exps = exps
        self.pair_binary = pair_binary
        self.valid_counts = sum(pair_binary.values())
        table = np.zeros((nb, nb))
        for (i, j), (r, _s) in slopes.items():
            if r > 0.6:
                table[i, j] = table[j, i] = 1
        self.connectivity = all(not np.any(table.diagonal(i) == 0) for i in range(1, min(nb, 3)))

    @property
    def hcm(self):
        """Heterozygosity-corrected upper-triangular matrix (assembly.py:557)."""
        if self._hcm is None:
            self._hcm = self._device_assembly().hcm()
        return self._hcm

    @property
    def symm_hcm(self):
        h = self.hcm.copy()
        x, y = h.nonzero()
        h[y, x] = h[x, y]
        return h

    # ------------------------------------------------------------------ coordinates
    def index_to_coordinates(self, loop_list):
        """Gap-free indices -> {(c1, s1, e1, c2, s2, e2): [distance on the assembly, neo flag]}
        (assembly.py:574-596)."""
        lines = {}
        res = self.res
        for i, j, _v in loop_list:
            x, y = self.index_map[i], self.index_map[j]
            a, b = self.reverse[x], self.reverse[y]
            if a > b:
                a, b = b, a
            key = (a[0], a[1], a[1] + res, b[0], b[1], b[1] + res)
            lines[key] = [(y - x) * res, 0 if self.chains[x] == self.chains[y] else 1]
        return lines


# ----------------------------------------------------------------------------------------------
# assemble-complexSVs: SV filter, SV graph, allele chains (neoloop/assembly.py:12-275)
# ----------------------------------------------------------------------------------------------
def filterSV(clr, c1, c2, p1, p2, s1, s2, note, span, col, protocol, cutoff, expected_values):
    """One SV: keep it when the contacts across the junction decay like the global expected
    (R^2 above ``cutoff``) and are strong enough relative to a flank; returns the graph node and the two
    flanks trimmed to the detected bounds, or None (assembly.py:12-27)."""
    from .util import map_coordinates
    fu = Fusion(clr, c1, c2, p1, p2, s1, s2, note, span=span, col=col, protocol=protocol, trim=False,
                expected_values=expected_values)
    fu.detect_bounds()
    fu.allele_slope()
    fu.correct_heterozygosity()
    if not (fu.m_g_r > cutoff and max(fu.u_s, fu.d_s) > 0.1):
        return None
    where = map_coordinates(fu.k_p, clr.binsize, c1, c2)[1]
    up, down = where[fu.up_i], where[fu.down_i]
    return [(c1, p1, s1, c2, p2, s2, note), (c1, up[1], fu.k_p[1]), (c2, fu.k_p[2], down[1])]


def filterAssembly(clr, ID, span, col, path, protocol, expected):
    """A chain of SVs is connectable when every block pair passes its regression and neighbouring
    blocks are linked (assembly.py:29-36)."""
    wk = complexSV(clr, ID, span, col, protocol=protocol, expected_values=expected)
    wk.reorganize_matrix()
    wk.correct_heterozygosity()
    ok = (len(wk.pair_binary) == wk.valid_counts) and wk.connectivity
    if wk._asm is not None:
        wk._asm.close()
    return ok, path


def _flip(node):
    """The same SV read from its other end."""
    return node[3:6] + node[:3] + (node[-1],)


class assembleSV(object):
    """SV list -> filtered SVs -> directed graph of SVs whose flanks overlap -> maximal chains that
    hold together in the contact map ("alleles") (assembly.py:39-275)."""

    def __init__(self, clr, sv_fil, span=5000000, col='sweight', minIntra=500000,
                 n_jobs=1, protocol='insitu', r_cutoff=0.6, expected_values=None):
        from .util import load_translocation_fil
        self.clr = clr
        self.res = clr.binsize
        self.protocol = protocol
        self.span = span
        self.balance_type = col
        self.n_jobs = n_jobs
        self.expected = expected_values
        self.queue, self.lookup = [], {}
        for c1, p1, s1, c2, p2, s2, note in load_translocation_fil(sv_fil, clr.binsize, minIntra):
            rec = filterSV(clr, c1, c2, p1, p2, s1, s2, note, span, col, protocol, r_cutoff, expected_values)
            if rec is not None:
                self.queue.append(rec)
                self.lookup[rec[0]] = [rec[1], rec[2]]

    # -- graph ---------------------------------------------------------------------------------
    def build_graph(self):
        import networkx as nx
        nodes = {}
        for n, u, d in self.queue:
            nodes[(n, u, d)] = n                 # as listed ...
            nodes[(_flip(n), d, u)] = n          # ... and read from the other end
        G = nx.DiGraph()
        for a in nodes:
            for b in nodes:
                if nodes[a] == nodes[b]:
                    continue
                dis = self.connect(a, b)
                if dis < np.inf:
                    G.add_weighted_edges_from([(a, b, dis)])
        self.graph = G
        self.nodemap = nodes

    def _check_overlap(self, i1, i2):
        """(intersection / union, intersection) of two intervals (chrom, a, b) given in any order."""
        if i1[0] != i2[0]:
            return 0, 0
        a0, a1 = sorted(i1[1:])
        b0, b1 = sorted(i2[1:])
        if a1 <= b0 or b1 <= a0:
            return 0, 0
        pts = sorted((a0, a1, b0, b1))
        return (pts[2] - pts[1]) / (pts[3] - pts[0]), pts[2] - pts[1]

    def connect(self, r1, r2):
        """Edge weight r1 -> r2: the tail flank of r1 and the head flank of r2 must face each other
        ('-' then '+', or '+' then '-'), overlap by more than four bins, and the two-SV chain must pass
        filterAssembly; weight = union / intersection, else inf (assembly.py:190-223)."""
        (n1, _u1, d1), (n2, u2, _d2) = r1, r2
        if (n1[5] == '-') != (n2[2] == '+'):
            return np.inf
        o, itr = self._check_overlap(d1, u2)
        if not (o > 0 and itr > 4 * self.res):
            return np.inf
        p = (r1, r2)
        ok, _p = filterAssembly(self.clr, self._change_format(p), self.span, self.balance_type, p, self.protocol,
                                self.expected)
        return 1 / o if ok else np.inf

    def _containloop(self, path):
        """Pairs of non-adjacent flanks of a chain that overlap by more than 5 % (a chain that visits
        the same region twice)."""
        chain = [flank for p in path for flank in (p[1], p[2])]
        hits = []
        for i in range(len(chain)):
            for j in range(i + 2, len(chain)):
                if self._check_overlap(chain[i], chain[j])[0] > 0.05:
                    hits.append((chain[i], chain[j]))
        return hits

    def _getreverse(self, path):
        return [(_flip(n), d, u) for n, u, d in path[::-1]]

    def _issubset(self, p1, p2):
        """p1 occurs in p2 as a contiguous, equally ordered run."""
        pos = []
        for item in p1:
            if item not in p2:
                return False
            pos.append(p2.index(item))
        return bool(np.all(np.diff(np.r_[pos]) == 1))

    def find_complexSV(self):
        from networkx.algorithms import all_shortest_paths, has_path
        pool = []
        nodes = self.graph.nodes()
        for n1 in nodes:
            for n2 in nodes:
                if n1 == n2 or not has_path(self.graph, n1, n2):
                    continue
                for p in all_shortest_paths(self.graph, n1, n2, weight=None):     # edge weights are not used
                    if not self._containloop(p):
                        pool.append((len(p), p))
        pool.sort(reverse=True)
        log.info('Filtering %d redundant candidates ...', len(pool))
        held = [p for _len, p in pool
                if filterAssembly(self.clr, self._change_format(p), self.span, self.balance_type, p, self.protocol,
                                  self.expected)[0]]
        alleles = []
        for p in held:              # longest first: drop chains contained in a kept one, in either reading direction
            if not any(self._issubset(p, v) or self._issubset(p, self._getreverse(v)) for v in alleles):
                alleles.append(p)
        self.alleles = alleles

    # -- output ----------------------------------------------------------------------------------
    def _change_format(self, path):
        """Assembly line body: 'type,c1,p1,s1,c2,p2,s2' per SV, then the two outer bounds
        (assembly.py:225-252)."""
        names = [','.join(map(str, (p[0][-1],) + p[0][:-1])) for p in path]
        b1, b2 = path[0][1], path[-1][2]
        first = min(b1[1:]) if path[0][0][2] == '+' else max(b1[1:])
        last = max(b2[1:]) if path[-1][0][5] == '-' else min(b2[1:])
        return '\t'.join(names + ['%s,%s' % (b1[0], first), '%s,%s' % (b2[0], last)])

    def print(self, complexSVs, outfil):
        with open(outfil, 'w') as out:
            for path in complexSVs:
                out.write(self._change_format(path) + '\n')

    def output_all(self, outfil):
        with open(outfil, 'w') as out:
            for i, path in enumerate(self.alleles):
                out.write('A%d\t%s\n' % (i, self._change_format(path)))
            for i, sv in enumerate(self.lookup):
                path = [(sv, self.lookup[sv][0], self.lookup[sv][1])]
                out.write('C%d\t%s\n' % (i, self._change_format(path)))
