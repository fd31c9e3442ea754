"""scratch: config-5 six-frame scan alone (count + emit), timed with CUDA events."""
import ctypes, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))   # repo root
import torch
from magot_b200 import _lib, engine, synth
lib = _lib.lib
GENOME_BP = int(os.environ.get("SIX_BP", 3_100_000_000)); SEED = 4
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
n_orf, n_bytes = ctypes.c_int64(0), ctypes.c_int64(0)
reps = int(os.environ.get("SIX_REPS", 3))
for it in range(reps):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(stream)
    _lib.check(lib.mg_sixframe_count(g.handle, 0, len(layout), 100, ctypes.byref(n_orf), ctypes.byref(n_bytes), sp))
    e[1].record(stream)
    aa = torch.empty((n_bytes.value + 31) // 32 * 32 + 32, dtype=torch.uint8, device=dev)
    ea = torch.cuda.Event(enable_timing=True); ea.record(stream)
    _lib.check(lib.mg_sixframe_emit_device(g.handle, ctypes.c_void_p(aa.data_ptr()), None, sp))
    e[2].record(stream); torch.cuda.synchronize()
    print("orfs %d bytes %d count %.3f ms emit %.3f ms cksum %d" % (n_orf.value, n_bytes.value, e[0].elapsed_time(e[1]), ea.elapsed_time(e[2]), int(aa[:n_bytes.value].to(torch.int64).sum().item())))
