#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload): synthetic 3.1 Gbp human-scale genome (24 chromosomes + 170 scaffolds,
5 % N, 50 % soft-masked) resident on the GPU as two strand planes of 0.5 B/base each, and a GTF-shaped annotation of 200k
transcripts per GPU (weak scaling: every rank owns its own 200k-transcript batch over a replicated
genome; the path has no cross-shard exchange, so there is no collective on the data path).
One STEP = one pass of the hot path over one batch, producing all three products of config 4:
CDS nucleotide FASTA, CDS protein FASTA, exon-based transcript FASTA (K1 clamp+scan, K2 splice+RC,
K3 translate, FASTA framing on the device).
metric  spliced+translated Gbp/s = (CDS bp spliced + CDS bp translated + exon bp spliced) / time.
value   device-resident: interval tables already in HBM, outputs left in HBM (CUDA events).
e2e     through the C ABI with HOST buffers: pinned SoA tables -> H2D -> kernels -> D2H of the three
        texts into pinned memory, every step.
The reference arm (--impl reference) times the reference's own Python get_fasta (oracle/_ref, the shimmed
reference; else the oracle port) on the host cores on a bounded sample of the same workload shape.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GENOME_BP = int(os.environ.get("MAGOT_BENCH_GENOME_BP", 3_100_000_000))
N_TX = int(os.environ.get("MAGOT_BENCH_TX", 200_000))
SAMPLE_BP = int(os.environ.get("MAGOT_BENCH_SAMPLE_BP", 20_000_000))
SAMPLE_TX = int(os.environ.get("MAGOT_BENCH_SAMPLE_TX", 1290))
SEED = 4
METRIC = "spliced+translated Gbp/s"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "MEASURED_PEAKS.json (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# CPU legs (rank 0 only): the reference's own code on a bounded sample
# ------------------------------------------------------------------------------------------------------
def build_sample():
    """Scaled twin of the workload (same generators): host genome + annotation + GTF text."""
    from magot_b200 import synth
    layout = synth.contig_layout("human", SAMPLE_BP, SEED)
    contigs = synth.synth_genome_host(layout, SEED)
    ann = synth.synth_annotation(layout, SAMPLE_TX, SEED)
    return layout, contigs, ann


def _sample_bp(ann):
    return 2 * ann.spliced_bp("cds") + ann.spliced_bp("exon")


_REF_STATE = {}


def _ref_worker(args):
    which, seq_type, keys = args
    g = _REF_STATE[which]
    n = 0
    for k in keys:
        n += len(g.annotations.gene[k].get_fasta(seq_type=seq_type))
    return n


def reference_setup(layout, contigs, ann):
    """Build the reference's (or the port's) objects for the sample. Returns (kind, run_step(pool) -> bytes)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    names = [n for n, _ in layout]
    fasta = "".join(">%s\n%s\n" % (n, c.tobytes().decode("latin-1")) for n, c in zip(names, contigs))
    try:
        import ref_runner
        ref = ref_runner.ref()
    except Exception:
        ref = None
    if ref is not None:
        kind = "reference"
        # The reference's own classes, populated directly (what its read_gff builds for a GTF: gene ->
        # transcript -> CDS|exon); read_gff itself costs ~1 ms/line in the reference and is not the timed path.
        for which in ("cds", "exon"):
            gs = ref.GenomeSequence(fasta)
            g = ref.Genome(gs)
            aset = ref.AnnotationSet()
            aset.genome = g
            g.annotations = aset
            off, st, en = (ann.cds_off, ann.cds_start, ann.cds_end) if which == "cds" else (ann.exon_off, ann.exon_start, ann.exon_end)
            ftype = "CDS" if which == "cds" else "exon"
            if ftype not in aset.__dict__:
                aset.__dict__[ftype] = {}
            for t in range(ann.n_tx):
                tx = ann.names[t]
                gid = "g%d" % ann.gene_of[t]
                ctg = names[ann.contig[t]]
                sd = "-" if ann.strand[t] else "+"
                if gid not in aset.gene:
                    aset.gene[gid] = ref.ParentAnnotation(gid, ctg, "gene", [], None, sd, aset)
                aset.gene[gid].child_list.append(tx)
                kids = []
                for k in range(off[t], off[t + 1]):
                    cid = "%s-%s%d" % (tx, ftype, k - off[t])
                    kids.append(cid)
                    aset.__dict__[ftype][cid] = ref.BaseAnnotation(cid, ctg, (int(st[k]), int(en[k])), ftype, tx, sd, {}, aset)
                aset.transcript[tx] = ref.ParentAnnotation(tx, ctg, "transcript", kids, gid, sd, aset)
            _REF_STATE[which] = g
        keys = list(_REF_STATE["cds"].annotations.gene)
    else:
        kind = "port"
        import magot_oracle as mo
        gtf = ann.to_gtf(names)
        seqs, _ = mo.read_fasta(fasta)

        class _G(object):
            pass
        for which, kw in (("cds", {}), ("exon", {"base_features": ('exon', 'match_part', 'similarity', 'region'), "features_to_ignore": ('CDS',)})):
            aset = mo.read_gff(gtf, **kw)
            aset.genome = seqs
            g = _G()
            g.annotations = _G()

            class _Wrap(object):
                def __init__(self, o):
                    self.o = o

                def get_fasta(self, seq_type="nucleotide"):
                    return mo.parent_get_fasta(self.o, seq_type=seq_type)
            g.annotations.gene = {k: _Wrap(v) for k, v in aset.table('gene').items()}
            _REF_STATE[which] = g
        keys = list(_REF_STATE["cds"].annotations.gene)
    return kind, keys


def reference_step(pool, keys, workers):
    tasks = []
    for which, seq_type in (("cds", "nucleotide"), ("cds", "protein"), ("exon", "nucleotide")):
        chunk = max(1, (len(keys) + workers * 4 - 1) // (workers * 4))
        for i in range(0, len(keys), chunk):
            tasks.append((which, seq_type, keys[i:i + chunk]))
    if pool is None:
        return sum(_ref_worker(t) for t in tasks)
    return sum(pool.map(_ref_worker, tasks))


def cpu_reference_rate(steps, warmup, workers):
    """Gbp/s of the reference's get_fasta path on the sample with `workers` processes."""
    import multiprocessing as mp
    layout, contigs, ann = build_sample()
    kind, keys = reference_setup(layout, contigs, ann)
    bp = _sample_bp(ann)
    pool = mp.get_context("fork").Pool(workers) if workers > 1 else None
    try:
        for _ in range(warmup):
            reference_step(pool, keys, workers)
        t0 = time.perf_counter()
        for _ in range(steps):
            reference_step(pool, keys, workers)
        dt = (time.perf_counter() - t0) / max(steps, 1)
    finally:
        if pool is not None:
            pool.terminate()
    sample = "config-4 generators at 1/%d scale: %d bp genome, %d transcripts, CDS nuc + CDS protein + exon transcripts via get_fasta (%s, CPython %d.%d; Python 2.7 is not available)" % (
        max(1, GENOME_BP // SAMPLE_BP), SAMPLE_BP, ann.n_tx, "shimmed reference oracle/_ref" if kind == "reference" else "oracle port", sys.version_info[0], sys.version_info[1])
    return bp / dt / 1e9, dt, kind, sample


def host_phases():
    """Host-side annotation parsing (outside every timed region): the product's read_gff and the reference's own on GTF text of
    the sample (the reference is ~1 ms per line, so it gets the first 2 000 lines only)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from magot_b200 import genome as mg_genome
    layout, contigs, ann = build_sample()
    gtf = ann.to_gtf([n for n, _ in layout])
    n_lines = gtf.count("\n")
    t0 = time.perf_counter()
    mg_genome.read_gff(gtf)
    t_prod = time.perf_counter() - t0
    out = {"gtf_lines": n_lines, "product_read_gff_lines_per_s": round(n_lines / t_prod)}
    try:
        import contextlib
        import io
        import ref_runner
        ref = ref_runner.ref()
        head = "\n".join(gtf.split("\n")[:2000]) + "\n"
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            ref.read_gff(head)
        out["reference_read_gff_lines_per_s"] = round(2000 / (time.perf_counter() - t0))
    except Exception as e:
        out["reference_read_gff_error"] = str(e)[:200]
    return out


def fasta_ingest_rates(device):
    """FASTA ingest (outside every timed region): a 256 Mbp record with 60-column lines through
    mg_genome_pack_fasta (raw bytes to the device, line ends dropped there by K0f, packed by K0) next to what the host would
    spend only stripping the line ends (bytes.translate) and to the reference's own GenomeSequence loader (genome.py:856-877)
    on a 1/16 sample."""
    import numpy as np
    from magot_b200 import engine
    n = 256 << 20
    rng = np.random.default_rng(3)
    rows = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (n // 64, 61), dtype=np.uint8)]
    rows[:, 60] = 10
    raw = rows.reshape(-1)
    bases = (n // 64) * 60
    g = engine.DeviceGenome([bases], device=device)
    t0 = time.perf_counter()
    g.pack_fasta(0, raw)                                           # first pass: allocates the device buffers
    t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    g.pack_fasta(0, raw)                                           # second pass: buffers exist, pages are warm
    t_dev2 = time.perf_counter() - t0
    g.finalize()
    assert g.fetch(0, bases - 120, bases) == raw[-122:].tobytes().translate(None, b"\n")
    g.close()
    data = raw.tobytes()
    t0 = time.perf_counter()
    data.translate(None, b"\r\n")
    t_host = time.perf_counter() - t0
    out = {"raw_bytes": int(raw.size), "pack_fasta_GBps_first": round(raw.size / t_dev / 1e9, 2),
           "pack_fasta_GBps": round(raw.size / t_dev2 / 1e9, 2), "host_strip_only_GBps": round(raw.size / t_host / 1e9, 2)}
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_runner
        ref = ref_runner.ref()
        sample = ">c\n" + data[:raw.size // 16].decode("latin-1")
        t0 = time.perf_counter()
        ref.GenomeSequence(sample)
        out["reference_GenomeSequence_GBps"] = round(len(sample) / (time.perf_counter() - t0) / 1e9, 4)
    except Exception as e:
        out["reference_error"] = str(e)[:200]
    return out


def cpu_port_rate():
    """Single-threaded C restatement (oracle/oracle.c) on a larger sample: tight-loop CPU figure."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import coracle
    from magot_b200 import synth
    layout = synth.contig_layout("human", 60_000_000, SEED)
    contigs = synth.synth_genome_host(layout, SEED)
    ann = synth.synth_annotation(layout, 20_000, SEED)
    lens = np.array([a.size for a in contigs])
    raw = [a.tobytes() for a in contigs]
    t0 = time.perf_counter()
    bp = 0
    for which in ("cds", "exon"):
        tbl = ann.table(which, framing=False)
        lo = np.clip(tbl.seg_start - 1, 0, lens[tbl.seg_contig])
        hi = np.clip(tbl.seg_end, 0, lens[tbl.seg_contig])
        nuc, off = coracle.splice(raw, tbl.rec_seg_off, tbl.seg_contig, lo, hi, tbl.seg_strand)
        bp += nuc.size
        if which == "cds":
            coracle.splice_translate(nuc, off)
            bp += nuc.size
    dt = time.perf_counter() - t0
    return bp / dt / 1e9


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def gpu_arm(args):
    import numpy as np
    import torch
    from magot_b200 import _lib, engine, synth
    lib = _lib.lib
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.require_device(local)
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)

    # ---- resident genome: synthesised on the device (torch, plumbing) and packed by K0
    t_setup = time.perf_counter()
    layout = synth.contig_layout("human", GENOME_BP, SEED)
    g = engine.DeviceGenome([l for _, l in layout], device=local)
    t_pack = 0.0
    CH = 256 << 20
    for ci, (_, L) in enumerate(layout):
        for off in range(0, L, CH):
            n = min(CH, L - off)
            a = synth.synth_contig_device(n, SEED * 1000003 + ci * 64 + off // CH, dev)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            g.pack_device(ci, a.data_ptr(), n, offset=off, stream=sp)
            torch.cuda.synchronize()
            t_pack += time.perf_counter() - t0
            del a
    g.finalize()
    torch.cuda.empty_cache()

    # ---- this rank's batch: 200k transcripts (weak scaling: a different batch per rank)
    ann = synth.synth_annotation(layout, N_TX, SEED + 1000 * rank)
    tables = {"cds": ann.table("cds"), "exon": ann.table("exon")}
    S_cds, S_exon = ann.spliced_bp("cds"), ann.spliced_bp("exon")
    bp_step = 2 * S_cds + S_exon
    setup_s = time.perf_counter() - t_setup

    # pinned copies of the host tables for the e2e path
    def pin(a):
        t = torch.from_numpy(np.array(a, copy=True)).pin_memory()
        return t
    pinned = {}
    for k, t in tables.items():
        pinned[k] = {f: pin(getattr(t, f)) for f in ("rec_seg_off", "seg_contig", "seg_start", "seg_end", "seg_strand", "rec_lit_off",
                                                       "rec_pre_len", "rec_suf_len", "lit")}

    def P(t):
        return ctypes.c_void_p(t.data_ptr())

    def create_plan(k):
        p = pinned[k]
        h = ctypes.c_void_p()
        _lib.check(lib.mg_plan_create(g.handle, tables[k].n_rec, P(p["rec_seg_off"]), tables[k].n_seg, P(p["seg_contig"]), P(p["seg_start"]),
                                      P(p["seg_end"]), P(p["seg_strand"]), P(p["rec_lit_off"]), P(p["rec_pre_len"]), P(p["rec_suf_len"]),
                                      P(p["lit"]), p["lit"].numel(), None, sp, ctypes.byref(h)))
        return h

    def prepare(h):
        a, b = ctypes.c_int64(0), ctypes.c_int64(0)
        _lib.check(lib.mg_plan_prepare(h, _lib.MG_PROT_TRIMX, ctypes.byref(a), ctypes.byref(b), sp))
        return a.value, b.value

    h2d_bytes = sum(t.numel() * t.element_size() for k in pinned for t in pinned[k].values())

    # ---- device-resident plans + output buffers
    plans = {k: create_plan(k) for k in tables}
    sizes = {k: prepare(plans[k]) for k in tables}
    def _cap(k, j):                                  # host-side upper bound of a text size, see `caps` below
        t = tables[k]
        pay = t.approx_bytes_per_record() - t.rec_pre_len - t.rec_suf_len
        lit_bytes = int(t.rec_pre_len.astype(np.int64).sum() + t.rec_suf_len.astype(np.int64).sum())
        return (int(pay.sum()) + lit_bytes, int((pay // 3).sum()) + lit_bytes)[j]
    out_cds_n = torch.empty((_cap("cds", 0) + 31) // 32 * 32 + 32, dtype=torch.uint8, device=dev)
    out_cds_p = torch.empty((_cap("cds", 1) + 31) // 32 * 32 + 32, dtype=torch.uint8, device=dev)
    out_exon_n = torch.empty((_cap("exon", 0) + 31) // 32 * 32 + 32, dtype=torch.uint8, device=dev)
    d2h_bytes = sizes["cds"][0] + sizes["cds"][1] + sizes["exon"][0]
    ev = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731
    nuc_events = []

    step_events = []
    # The two plans of a step are independent: the CDS plan runs on `stream`, the exon plan on `stream_b`, so the
    # latency-bound plan kernels (and the host round trip for the totals) of one overlap the emit kernels of the other.
    stream_b = torch.cuda.Stream(device=dev)
    spb = ctypes.c_void_p(stream_b.cuda_stream)

    # Upper bounds of the text sizes from the HOST tables (sum of end-start+1 per segment + framing; a third of it for the
    # protein): with them K1 needs no host round trip (mg_plan_prepare_async) and K1 -> K2 -> K3 queue back to back.
    caps = {}
    for k, t in tables.items():
        pay = t.approx_bytes_per_record() - t.rec_pre_len - t.rec_suf_len
        lit_bytes = int(t.rec_pre_len.astype(np.int64).sum() + t.rec_suf_len.astype(np.int64).sum())
        caps[k] = (int(pay.sum()) + lit_bytes, int((pay // 3).sum()) + lit_bytes)
        assert caps[k][0] >= sizes[k][0] and caps[k][1] >= sizes[k][1]

    def prepare_on(h, s, k):
        _lib.check(lib.mg_plan_prepare_async(h, _lib.MG_PROT_TRIMX, caps[k][0], caps[k][1], s))

    last_heavy = [None]                              # event after the previous step's exon emit

    def device_step(record):
        # K1 of each plan is queued first and floats; the bandwidth-heavy emit kernels of the two plans are put in series
        # with events (K2 cds -> K3 cds -> K2 exon -> next step's K2 cds), so that a K1 overlaps the other plan's emits but
        # two emit kernels never share the GPU (and the per-kernel event times below stay those of the kernel alone).
        # (K1 on high-priority streams of its own was measured: same step time, but it slows the emit kernel it overlaps.)
        ea = ev()
        ea.record(stream)
        prepare_on(plans["cds"], sp, "cds")
        e0, e1, ep = ev(), ev(), ev()
        if last_heavy[0] is not None:
            stream.wait_event(last_heavy[0])
        e0.record(stream)
        _lib.check(lib.mg_emit_nuc_device(plans["cds"], P(out_cds_n), sp))
        e1.record(stream)
        _lib.check(lib.mg_emit_prot_device(plans["cds"], P(out_cds_p), sp))
        ep.record(stream)
        eb = ev()
        eb.record(stream_b)
        prepare_on(plans["exon"], spb, "exon")
        e2, e3 = ev(), ev()
        stream_b.wait_event(ep)
        e2.record(stream_b)
        _lib.check(lib.mg_emit_nuc_device(plans["exon"], P(out_exon_n), spb))
        e3.record(stream_b)
        last_heavy[0] = e3
        if record:
            nuc_events.append((e0, e1, e2, e3))
            step_events.append((ea, e0, e1, ep, eb, e2, e3))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        device_step(False)
    clocks = ClockSampler(local)
    clocks.start()
    barrier()
    l0 = lib.mg_kernel_launches()
    s_ev, e_ev = ev(), ev()
    s_ev.record(stream)
    for _ in range(args.steps):
        device_step(True)
    stream.wait_stream(stream_b)
    e_ev.record(stream)
    barrier()
    launches = lib.mg_kernel_launches() - l0
    dev_ms = s_ev.elapsed_time(e_ev) / args.steps
    for k in tables:                                 # the sizes the device found in the last step are the ones of the sync prepare
        a, b = ctypes.c_int64(0), ctypes.c_int64(0)
        _lib.check(lib.mg_plan_totals(plans[k], ctypes.byref(a), ctypes.byref(b), sp if k == "cds" else spb))
        assert (a.value, b.value) == tuple(sizes[k]), (k, a.value, b.value, sizes[k])
    nuc_ms_cds = sum(a.elapsed_time(b) for a, b, _, _ in nuc_events) / len(nuc_events)
    nuc_ms_exon = sum(c.elapsed_time(d) for _, _, c, d in nuc_events) / len(nuc_events)
    _seg = lambda i: sum(t[i].elapsed_time(t[i + 1]) for t in step_events) / len(step_events)   # noqa: E731
    breakdown = {"k1_plan_cds_ms": round(_seg(0), 4), "k2_nuc_cds_ms": round(_seg(1), 4), "k3_prot_cds_ms": round(_seg(2), 4),
                 "k1_plan_exon_ms": round(_seg(4), 4), "k2_nuc_exon_ms": round(_seg(5), 4),
                 "note": "CDS plan on one stream, exon plan on a second; k1_* include waiting for the other plan's emit kernels (the emit "
                         "kernels are serialised by events, the plan kernels overlap them), so the segments sum to more than ms_per_step"}

    # ---- end to end through the C ABI with host buffers
    # Successive batches are double-buffered, the way a caller streaming batches would do it: step i runs on stream pair
    # i % 2 into host buffer set i % 2 and is only waited for (and its plans destroyed) when step i + 2 needs the pair
    # again, so the device->host copies of one step (what bounds the step: 674 MB over PCIe) overlap the table upload
    # and the kernels of the next.  Every step still uploads its own interval tables from pinned host memory and
    # copies its three texts back to the host inside the timed region.
    host_sets = [(torch.empty(sizes["cds"][0], dtype=torch.uint8, pin_memory=True),
                  torch.empty(sizes["cds"][1], dtype=torch.uint8, pin_memory=True),
                  torch.empty(sizes["exon"][0], dtype=torch.uint8, pin_memory=True)) for _ in range(2)]
    pairs = [(stream, stream_b), (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))]

    def create_plan_on(k, s):
        p = pinned[k]
        h = ctypes.c_void_p()
        _lib.check(lib.mg_plan_create(g.handle, tables[k].n_rec, P(p["rec_seg_off"]), tables[k].n_seg, P(p["seg_contig"]), P(p["seg_start"]),
                                      P(p["seg_end"]), P(p["seg_strand"]), P(p["rec_lit_off"]), P(p["rec_pre_len"]), P(p["rec_suf_len"]),
                                      P(p["lit"]), p["lit"].numel(), None, s, ctypes.byref(h)))
        return h

    in_flight = [None, None]

    def e2e_retire(slot):
        if in_flight[slot] is not None:
            hc, he, sa, sb = in_flight[slot]
            _lib.check(lib.mg_stream_sync(local, sa))
            _lib.check(lib.mg_stream_sync(local, sb))
            lib.mg_plan_destroy(hc)
            lib.mg_plan_destroy(he)
            in_flight[slot] = None

    def e2e_step(i):
        slot = i & 1
        e2e_retire(slot)
        sa = ctypes.c_void_p(pairs[slot][0].cuda_stream)
        sb = ctypes.c_void_p(pairs[slot][1].cuda_stream)
        h_cds_n, h_cds_p, h_exon_n = host_sets[slot]
        hc = create_plan_on("cds", sa)
        he = create_plan_on("exon", sb)
        a, b = ctypes.c_int64(0), ctypes.c_int64(0)
        _lib.check(lib.mg_plan_prepare(hc, _lib.MG_PROT_TRIMX, ctypes.byref(a), ctypes.byref(b), sa))
        _lib.check(lib.mg_emit_nuc_host(hc, P(h_cds_n), sa))
        _lib.check(lib.mg_emit_prot_host(hc, P(h_cds_p), sa))
        _lib.check(lib.mg_plan_prepare(he, _lib.MG_PROT_TRIMX, ctypes.byref(a), ctypes.byref(b), sb))
        _lib.check(lib.mg_emit_nuc_host(he, P(h_exon_n), sb))
        in_flight[slot] = (hc, he, sa, sb)

    for i in range(max(2, min(args.warmup, 4))):
        e2e_step(i)
    e2e_retire(0)
    e2e_retire(1)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    e2e_retire(0)
    e2e_retire(1)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps      # host clock around fully synchronised work on four streams
    barrier()
    # both host buffer sets hold the same texts as the device-resident step produced
    for hs in host_sets:
        for h_t, d_t, n in ((hs[0], out_cds_n, sizes["cds"][0]), (hs[1], out_cds_p, sizes["cds"][1]), (hs[2], out_exon_n, sizes["exon"][0])):
            assert torch.equal(h_t, d_t[:n].cpu()), "end-to-end text differs from the device-resident text"
    clk = clocks.stop()

    # ---- config 5 (extra information, outside the timed steps): six-frame ORF scan of the whole genome, min ORF 100 aa
    six = None
    if not args.no_sixframe:
        try:
            # contigs are independent: each rank scans its LPT share of the contigs (longest first, each to the least loaded
            # GPU; strong scaling, no exchange step); time = max over ranks, ORFs and residues summed
            from magot_b200 import orfs as _orfs
            shard = _orfs.lpt_shards([l for _, l in layout], world)[rank]
            ids = np.ascontiguousarray(shard, dtype=np.int64)
            my_bp = sum(layout[c][1] for c in shard)
            n_orf, n_bytes = ctypes.c_int64(0), ctypes.c_int64(0)

            def six_count():
                _lib.check(lib.mg_sixframe_count_list(g.handle, ids.size, ctypes.c_void_p(ids.ctypes.data), 100, ctypes.byref(n_orf),
                                                      ctypes.byref(n_bytes), sp))
            six_count()                                  # warm-up
            torch.cuda.synchronize()
            a0, a1, a2 = ev(), ev(), ev()
            a0.record(stream)
            six_count()
            a1.record(stream)
            aa_dev = torch.empty((n_bytes.value + 31) // 32 * 32 + 32, dtype=torch.uint8, device=dev)
            a1b = ev()
            a1b.record(stream)
            _lib.check(lib.mg_sixframe_emit_device(g.handle, P(aa_dev), None, sp))
            a2.record(stream)
            torch.cuda.synchronize()
            t_scan, t_emit = a0.elapsed_time(a1), a1b.elapsed_time(a2)
            tot_orf, tot_bytes, max_bp = n_orf.value, n_bytes.value, my_bp
            if dist is not None:
                tt = torch.tensor([t_scan, t_emit, float(my_bp)], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t_scan, t_emit, max_bp = [float(x) for x in tt.tolist()]
                ts = torch.tensor([n_orf.value, n_bytes.value], dtype=torch.float64, device=dev)
                dist.all_reduce(ts, op=dist.ReduceOp.SUM)
                tot_orf, tot_bytes = [int(x) for x in ts.tolist()]
            six = {"workload": "config 5: six-frame translation + ORF scan of the whole genome, min ORF 100 aa (reference semantics of Sequence.get_orfs)",
                   "sharding": "contigs assigned to %d GPU(s) longest-first to the least loaded (LPT); largest share %.3f Gbp; strong scaling, "
                               "times are the max over ranks" % (world, max_bp / 1e9),
                   "orfs": tot_orf, "aa_bytes": tot_bytes, "scan_ms": round(t_scan, 3), "emit_ms": round(t_emit, 3),
                   "genome_Gbp_per_s": round(GENOME_BP / ((t_scan + t_emit) * 1e-3) / 1e9, 1),
                   "six_frame_Gbp_per_s": round(6 * GENOME_BP / ((t_scan + t_emit) * 1e-3) / 1e9, 1),
                   "algorithmic_GBps": round((GENOME_BP * 0.5 + tot_bytes + 32 * tot_orf) / ((t_scan + t_emit) * 1e-3) / 1e9, 1),
                   "bound": "instruction issue and CTA barriers, not HBM (the genome is read once: 1.56 GB): see DESIGN.md section 4"}
            del aa_dev
        except Exception as e:
            six = {"error": str(e)[:300]}
            if dist is not None:
                raise

    # max over ranks
    if dist is not None:
        t = torch.tensor([dev_ms, e2e_ms, nuc_ms_cds, nuc_ms_exon], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, nuc_ms_cds, nuc_ms_exon = [float(x) for x in t.tolist()]
        tot = torch.tensor([bp_step, h2d_bytes, d2h_bytes, launches], dtype=torch.float64, device=dev)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        bp_all, h2d_all, d2h_all, launches_all = [float(x) for x in tot.tolist()]
    else:
        bp_all, h2d_all, d2h_all, launches_all = bp_step, h2d_bytes, d2h_bytes, launches

    # ---- roofline of the dominant kernel (k_emit_nuc), per launch, from this rank's tables
    peak, peak_src = peaks()

    def alg_bytes(S, tbl, total_text):
        # S*0.5 (packed read) + text written + 14 B/segment + 8 B/record (+ literal bytes read)
        return S * 0.5 + total_text + tbl.n_seg * 14 + tbl.n_rec * 8 + tbl.lit.size
    ab_cds = alg_bytes(S_cds, tables["cds"], sizes["cds"][0])
    ab_exon = alg_bytes(S_exon, tables["exon"], sizes["exon"][0])
    ach = (ab_cds + ab_exon) / ((nuc_ms_cds + nuc_ms_exon) * 1e-3) / 1e9
    roofline = {"kernel": "k_emit_nuc (K2 splice + per-segment RC + FASTA framing)", "bound": "hbm", "achieved": round(ach, 1),
                "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "traffic": 567e6 if (GENOME_BP == 3_100_000_000 and N_TX == 200_000) else None,
                "traffic_source": "dram__bytes_read+write per launch, mean of the CDS (241+192 MB) and exon (369+332 MB) launches, ncu --set full, profiles/r1ag_emit_plan_raw.csv",
                "peak_source": peak_src,
                "frac_of_nominal_8TBs": round(ach / 8000.0, 4),
                "launches_per_step": 2, "avg_launch_ms": round((nuc_ms_cds + nuc_ms_exon) / 2, 4),
                "algorithmic_bytes_per_launch": int((ab_cds + ab_exon) / 2),
                "per_launch": {"cds": {"ms": round(nuc_ms_cds, 4), "GBps": round(ab_cds / nuc_ms_cds / 1e6, 1)},
                               "exon": {"ms": round(nuc_ms_exon, 4), "GBps": round(ab_exon / nuc_ms_exon / 1e6, 1)}},
                "step_breakdown_ms": breakdown}

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                v, dt, kind, sample = cpu_reference_rate(1, 0, 1)
                cpu = {"value": v, "unit": "Gbp/s", "cores": 1, "kind": kind, "sample": sample, "seconds": round(dt, 2),
                       "host_cpus": os.cpu_count()}
                try:
                    cpu["c_port_single_thread_Gbps"] = round(cpu_port_rate(), 4)
                except Exception as e:      # the C port is extra information only
                    cpu["c_port_error"] = str(e)[:200]
                try:
                    cpu["host_annotation_parse"] = host_phases()
                except Exception as e:
                    cpu["host_annotation_parse"] = {"error": str(e)[:200]}
                try:
                    cpu["fasta_ingest"] = fasta_ingest_rates(local)
                except Exception as e:
                    cpu["fasta_ingest"] = {"error": str(e)[:200]}
            except Exception as e:
                cpu = {"value": None, "unit": "Gbp/s", "cores": 1, "kind": "port", "sample": "failed: %s" % str(e)[:300]}
        line = {
            "metric": METRIC, "value": bp_all / (dev_ms * 1e-3) / 1e9, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "config 4: synthetic %.2f Gbp human-scale genome (replicated per GPU; forward + reverse-complement planes, 0.5 B/base each) + %d-transcript GTF-shaped batch per GPU; products: CDS nucleotide FASTA + CDS protein FASTA + exon-based transcript FASTA" % (GENOME_BP / 1e9, N_TX),
                       "genome_bp": GENOME_BP, "transcripts_per_gpu": N_TX, "cds_segments": int(tables["cds"].n_seg),
                       "exons": int(tables["exon"].n_seg), "spliced_cds_bp": S_cds, "spliced_exon_bp": S_exon,
                       "bp_per_step_per_gpu": bp_step, "parallelism": "transcript batches per GPU, genome replicated, no collective; per GPU the CDS and exon plans run on two CUDA streams",
                       "l2": "no flush: each step streams ~%.1f GB of distinct output + genome lines, far above the 126 MB L2" % ((d2h_bytes + 0.5 * bp_step) / 1e9),
                       "genome_device_bytes": int(g.device_bytes()), "pack_s": round(t_pack, 3), "setup_s": round(setup_s, 1)},
            "clocks": clk,
            "e2e": {"value": bp_all / (e2e_ms * 1e-3) / 1e9, "unit": "Gbp/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all)},
            "gpu_launches": int(launches_all),
            "roofline": roofline,
        }
        if six is not None:
            line["sixframe"] = six
        if cpu is not None:
            line["cpu_baseline"] = cpu
    for h in plans.values():
        lib.mg_plan_destroy(h)
    g.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    workers = max(1, min(os.cpu_count() or 1, 32))
    v, dt, kind, sample = cpu_reference_rate(max(args.steps, 1), min(args.warmup, 1), workers)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Gbp/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "config 4 (bounded sample): each step = the reference's get_fasta over the sample, all host cores"},
            "cpu_baseline": {"value": v, "unit": "Gbp/s", "cores": workers, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="magot_b200", choices=["magot_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sixframe", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        args.warmup = max(args.warmup, 3)
        gpu_arm(args)


if __name__ == "__main__":
    main()
