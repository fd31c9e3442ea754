"""scratch: K2 per-launch times (CDS + exon launches of config 4) for the library named by MAGOT_B200_LIB / MAGOT_EMIT."""
import ctypes, sys, os, zlib, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))   # repo root
import numpy as np, torch
from magot_b200 import _lib, engine, synth
lib = _lib.lib
GENOME_BP = int(os.environ.get("MAGOT_BENCH_GENOME_BP", 3_100_000_000)); SEED = 4
N_TX = int(os.environ.get("MAGOT_BENCH_TX", 200_000))
tag = sys.argv[1] if len(sys.argv) > 1 else "?"
NIT = int(os.environ.get("K2AB_REPS", 12)); SKIP = min(3, NIT - 1)
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
res = {"tag": tag}
for which in ("cds", "exon"):
    t = ann.table(which)
    plan = engine.Plan(g, t); nuc, prot = plan.prepare()
    out = torch.empty((nuc + 31) // 32 * 32 + 32, dtype=torch.uint8, device=dev)
    times = []
    for it in range(NIT):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); _lib.check(lib.mg_emit_nuc_device(plan.handle, ctypes.c_void_p(out.data_ptr()), sp)); e1.record(stream)
        torch.cuda.synchronize(); times.append(e0.elapsed_time(e1))
    times = sorted(times[SKIP:])
    crc = zlib.crc32(out[:nuc].cpu().numpy().tobytes())
    alg = 0.5 * ann.spliced_bp(which) + nuc + t.n_seg * 14 + t.n_rec * 8 + t.lit.size
    res[which] = {"ms_min": round(times[0], 4), "ms_med": round(times[len(times) // 2], 4), "crc": crc, "bytes": nuc,
                  "alg_GBps_med": round(alg / times[len(times) // 2] / 1e6, 1), "frac": round(alg / times[len(times) // 2] / 1e6 / 6456.2, 4)}
    if which == "cds" and hasattr(lib, "mg_emit_nuc_prot_device"):
        outp = torch.empty((prot + 31) // 32 * 32 + 32, dtype=torch.uint8, device=dev)
        tt = []
        for it in range(NIT):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); _lib.check(lib.mg_emit_nuc_prot_device(plan.handle, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(outp.data_ptr()), sp)); e1.record(stream)
            torch.cuda.synchronize(); tt.append(e0.elapsed_time(e1))
        tt = sorted(tt[SKIP:])
        algf = alg + prot
        res["cds_fused"] = {"ms_min": round(tt[0], 4), "ms_med": round(tt[len(tt) // 2], 4), "crc_nuc": zlib.crc32(out[:nuc].cpu().numpy().tobytes()),
                            "crc_prot": zlib.crc32(outp[:prot].cpu().numpy().tobytes()), "frac": round(algf / tt[len(tt) // 2] / 1e6 / 6456.2, 4)}
        tp = []
        for it in range(NIT):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); _lib.check(lib.mg_emit_prot_device(plan.handle, ctypes.c_void_p(outp.data_ptr()), sp)); e1.record(stream)
            torch.cuda.synchronize(); tp.append(e0.elapsed_time(e1))
        tp = sorted(tp[SKIP:])
        res["cds_k3"] = {"ms_med": round(tp[len(tp) // 2], 4), "crc_prot": zlib.crc32(outp[:prot].cpu().numpy().tobytes())}
    plan.close(); del out
print(json.dumps(res))
