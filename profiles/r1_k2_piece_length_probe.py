"""scratch: K2 exon launch time vs piece length (is the kernel or the access pattern the limit?)"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from magot_b200 import _lib, engine, synth
lib = _lib.lib
GENOME_BP = 3_100_000_000; SEED = 4
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
for med, mean_ex, ntx in ((140, 9.5, 200_000), (600, 9.5, 50_000), (3000, 9.5, 10_000), (30000, 3.0, 3_000)):
    ann = synth.synth_annotation(layout, ntx, SEED, exon_median=med, mean_exons=mean_ex, intron_median=max(1000, med))
    t = ann.table("exon")
    plan = engine.Plan(g, t); nuc, _ = plan.prepare()
    out = torch.empty((nuc + 31) // 32 * 32, dtype=torch.uint8, device=dev)
    best = 1e9
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); _lib.check(lib.mg_emit_nuc_device(plan.handle, ctypes.c_void_p(out.data_ptr()), sp)); e1.record(stream)
        torch.cuda.synchronize(); ms = e0.elapsed_time(e1)
        if it >= 2: best = min(best, ms)
    alg = 0.5 * ann.spliced_bp("exon") + nuc + t.n_seg * 14 + t.n_rec * 8
    print("exon_median %6d: %d segs, %.0f MB out, %.4f ms, out %.0f GB/s, algorithmic %.0f GB/s (%.2f of 6456)" % (med, t.n_seg, nuc / 1e6, best, nuc / best / 1e6, alg / best / 1e6, alg / best / 1e6 / 6456.2))
    plan.close(); del out
