"""scratch: one merged launch (mg_emit_products_device) against the three single-product launches of config 4."""
import ctypes, sys, os, zlib, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))   # repo root
import numpy as np, torch
from magot_b200 import _lib, engine, synth
lib = _lib.lib
GENOME_BP = int(os.environ.get("MAGOT_BENCH_GENOME_BP", 3_100_000_000)); SEED = 4
N_TX = int(os.environ.get("MAGOT_BENCH_TX", 200_000))
LAGS = [int(x) for x in os.environ.get("LAGS", "0,20000,60000,100000,200000").split(",")]
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
stream = torch.cuda.current_stream(); sp = ctypes.c_void_p(stream.cuda_stream)
layout = synth.contig_layout("human", GENOME_BP, SEED)
g = engine.DeviceGenome([l for _, l in layout], device=0)
CH = 256 << 20
for ci, (_, L) in enumerate(layout):
    for off in range(0, L, CH):
        n = min(CH, L - off)
        a = synth.synth_contig_device(n, SEED * 1000003 + ci * 64 + off // CH, dev)
        g.pack_device(ci, a.data_ptr(), n, offset=off, stream=sp); torch.cuda.synchronize(); del a
g.finalize(); torch.cuda.empty_cache()
ann = synth.synth_annotation(layout, N_TX, SEED)
plans, outs, sizes, alg = {}, {}, {}, {}
for which in ("cds", "exon"):
    t = ann.table(which)
    plans[which] = engine.Plan(g, t); nuc, prot = plans[which].prepare()
    sizes[which] = (nuc, prot)
    outs[which + "_n"] = torch.zeros((nuc + 31) // 32 * 32 + 32, dtype=torch.uint8, device=dev)
    alg[which] = 0.5 * ann.spliced_bp(which) + nuc + t.n_seg * 14 + t.n_rec * 8 + t.lit.size
outs["cds_p"] = torch.zeros((sizes["cds"][1] + 31) // 32 * 32 + 32, dtype=torch.uint8, device=dev)
alg["prot"] = 0.5 * ann.spliced_bp("cds") + sizes["cds"][1] + ann.table("cds").n_seg * 14 + ann.table("cds").n_rec * 8
P = lambda t: ctypes.c_void_p(t.data_ptr())
REPS = int(os.environ.get("REPS", 12))
def timeit(fn, reps=REPS):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[min(3, reps - 1):]); return round(ts[len(ts) // 2], 4)
def crcs():
    return [zlib.crc32(outs[k][:n].cpu().numpy().tobytes()) for k, n in (("exon_n", sizes["exon"][0]), ("cds_n", sizes["cds"][0]), ("cds_p", sizes["cds"][1]))]
res = {}
hc, he = plans["cds"].handle, plans["exon"].handle
res["k2_exon"] = timeit(lambda: _lib.check(lib.mg_emit_nuc_device(he, P(outs["exon_n"]), sp)))
res["k2_cds"] = timeit(lambda: _lib.check(lib.mg_emit_nuc_device(hc, P(outs["cds_n"]), sp)))
res["k3"] = timeit(lambda: _lib.check(lib.mg_emit_prot_device(hc, P(outs["cds_p"]), sp)))
def three():
    _lib.check(lib.mg_emit_nuc_device(he, P(outs["exon_n"]), sp)); _lib.check(lib.mg_emit_nuc_device(hc, P(outs["cds_n"]), sp)); _lib.check(lib.mg_emit_prot_device(hc, P(outs["cds_p"]), sp))
res["three_in_series"] = timeit(three)
want = crcs()
tot_alg = alg["cds"] + alg["exon"] + alg["prot"]
res["alg_MB"] = {k: round(v / 1e6, 1) for k, v in alg.items()}
for lag in LAGS:
    _lib.check(lib.mg_tune(b"multi_lag", lag))
    for k in outs: outs[k].zero_()
    ms = timeit(lambda: _lib.check(lib.mg_emit_products_device(he, P(outs["exon_n"]), hc, P(outs["cds_n"]), P(outs["cds_p"]), sp)))
    ok = crcs() == want
    res["multi_lag%d" % lag] = {"ms": ms, "ok": ok, "frac": round(tot_alg / ms / 1e6 / 6456.2, 4)}
_lib.check(lib.mg_tune(b"multi_lag", LAGS[len(LAGS) // 2]))
for k in outs: outs[k].zero_()
ms = timeit(lambda: _lib.check(lib.mg_emit_products_device(he, P(outs["exon_n"]), hc, P(outs["cds_n"]), None, sp)))
res["multi_two_nuc"] = {"ms": ms, "ok": crcs()[:2] == want[:2], "frac": round((alg["cds"] + alg["exon"]) / ms / 1e6 / 6456.2, 4)}
ms = timeit(lambda: _lib.check(lib.mg_emit_products_device(None, None, hc, P(outs["cds_n"]), P(outs["cds_p"]), sp)))
res["multi_cds_only"] = {"ms": ms, "ok": crcs() == want, "frac": round((alg["cds"] + alg["prot"]) / ms / 1e6 / 6456.2, 4)}
print(json.dumps(res))
