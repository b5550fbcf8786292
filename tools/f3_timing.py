#!/usr/bin/env python
"""Row f-3 timing (gpurun): assembleSV on the fixture workload, through the cooler surface and from the resident
pixel table.  With --reference (build container only) the REAL reference runs the same workload on the CPU."""
import argparse, json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neoloopfinder_b200.coolshim import PixelCooler, SyntheticCooler

ap = argparse.ArgumentParser()
ap.add_argument("--reference", action="store_true")
args = ap.parse_args()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
z = json.load(open(os.path.join(root, "tests", "golden", "assemble_10k.json")))
clr = SyntheticCooler.from_spec(z["cooler_spec"])
sv_file = os.path.join(tempfile.mkdtemp(), "svs.txt")
open(sv_file, "w").write("\n".join(z["sv_lines"]) + "\n")
expected = {i: v for i, v in enumerate(z["expected"])}
out = {}


def run(cls, src):
    t0 = time.perf_counter()
    w = cls(src, sv_file, span=z["span"], col=z["col"], n_jobs=1, protocol="insitu", minIntra=z["min_intra"],
            expected_values=expected)
    t1 = time.perf_counter()
    w.build_graph()
    w.find_complexSV()
    t2 = time.perf_counter()
    return {"filter_svs_s": round(t1 - t0, 3), "graph_and_chains_s": round(t2 - t1, 3), "total_s": round(t2 - t0, 3),
            "kept": len(w.queue), "alleles": len(w.alleles)}


if args.reference:
    from oracle.run_reference import import_reference
    ra = import_reference()[1]
    out["reference_cpu"] = run(ra.assembleSV, clr)
else:
    from neoloopfinder_b200.assembly import assembleSV
    run(assembleSV, clr)                                   # warm-up (context, kernels)
    out["gpu_cooler_surface"] = run(assembleSV, clr)
    px = PixelCooler.from_cooler(clr)
    px.device_pixels(None, z["col"])
    out["gpu_resident_pixels"] = run(assembleSV, px)
print(json.dumps(out))
