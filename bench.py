#!/usr/bin/env python3
"""bench.py -- headline benchmark of the batched compression pipeline (BASELINE.json metric:
"compress/decompress GB/s, batched 512x512 RAW -m -a, 1/2/4/8 B200; bpc").

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, libhc_b200.so)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU codec on host cores
    python bench.py --workload c4|c5|c3m ...                 # the other BASELINE configs (not the default line)
    python bench.py --total-files 32768 ...                  # strong scaling: a fixed batch cut over the ranks

A "step" = one pass of the hot path over one batch: compress every file of the batch, then
decompress every .out again.  Default workload (config.workload) C3ma: 4096 synthetic 512x512
8-bit images per GPU (SURVEY 8d classes walk / smooth / random / const, seed 1234+i), -m -a -w 512.
Files are independent, so a batch shards over the ranks with no data-path collective; the only
exchange is the all-gather of the per-file output sizes (hc_shard_sizes_allgather: NCCL), from which
every rank derives the global offsets table.

value   = uncompressed bytes pushed through compress AND decompress by all ranks / device time (CUDA
          events, inputs resident in HBM, max over ranks).  Consecutive steps alternate over up to eight codec
          streams (config.overlap): FGK is bound by the latency of its longest stream, so the tail of one step's
          FGK kernels shares the GPU with the following steps (measured: 4 streams 10.9, 6: 11.2, 8: 11.4 GB/s).  `sequential_ms_per_step` = one stream.
e2e     = same metric through the asynchronous host API (hc_pipeline_*, depth 4) with pinned HOST
          buffers, host<->device copies inside the timed region, all K steps.
roofline= dominant kernel (FGK: instruction issue, see roofline.note) and, in `stages`, every
          transform kernel's achieved algorithmic GB/s against MEASURED_PEAKS.json:hbm_gbs.
cpu_baseline = the unmodified reference binary (oracle/_ref, built from /root/reference/src) run
          as one process per file on ALL host cores over a bounded stratified sample (longest files first).
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "huffman-codec_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import synth  # noqa: E402

METRIC = "compress+decompress throughput, batched 512x512 RAW -m -a"
UNIT = "GB/s"

# name -> (side, classes, seed0, use_adapt, default files per GPU, text)
WORKLOADS = {
    "c3ma": (512, synth.CLASSES, 1234, True, 4096, "C3ma: %d x (512x512 u8) synthetic per GPU (walk/smooth/random/const, seed 1234+i), -m -a -w 512"),
    "c3m": (512, synth.CLASSES, 1234, False, 4096, "C3m: %d x (512x512 u8) synthetic per GPU (walk/smooth/random/const, seed 1234+i), -m"),
    "c4": (4096, synth.CLASSES, 2000, True, 128, "C4: %d x (4096x4096 u8) synthetic per GPU (walk/smooth/random/const, seed 2000+i), -m -a -w 4096"),
    "c5": (512, ("random", "smooth", "const"), 3000, True, 16384,
           "C5: %d x (512x512 u8) per GPU in thirds random/smooth/const (seed 3000+i) + a Fibonacci-weighted file + a file of runs around 258, -m -a -w 512"),
}
_CLASSES_ENV = os.environ.get("HC_BENCH_CLASSES")       # experiments only: e.g. "random"


def _gen_one(args):
    i, seed0, side, classes, special = args
    if special and i == 0:
        return synth.image("fib", side, seed0).reshape(-1)
    if special and i == 1:
        return synth.image("longrun", side, seed0 + 1).reshape(-1)
    return synth.image(classes[i % len(classes)], side, seed0 + i).reshape(-1)


def make_batch(lo, hi, wl, procs):
    """files lo..hi of the workload (global indices): (hi - lo) x side^2 u8."""
    side, classes, seed0 = wl[0], wl[1], wl[2]
    if _CLASSES_ENV:
        classes = tuple(_CLASSES_ENV.split(","))
    special = wl is WORKLOADS["c5"]
    count = hi - lo
    out = np.empty((count, side * side), np.uint8)
    jobs = [(i, seed0, side, classes, special) for i in range(lo, hi)]
    if procs > 1 and count >= 32:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(procs, 32)) as pool:
            for k, a in enumerate(pool.imap(_gen_one, jobs, chunksize=4 if side > 512 else 16)):
                out[k] = a
    else:
        for k, j in enumerate(jobs):
            out[k] = _gen_one(j)
    return out, classes


# ------------------------------------------------------------------ reference arm (CPU)
def _ref_one(job):
    binary, flags, path = job
    t0 = time.perf_counter()
    r1 = subprocess.run([binary, "-c"] + flags + ["-i", path, "-o", path + ".out"], capture_output=True)
    r2 = subprocess.run([binary, "-d", "-i", path + ".out", "-o", path + ".dec"], capture_output=True)
    return r1.returncode, r2.returncode, time.perf_counter() - t0


def reference_sample(files, flags, cores, keep_outputs=False, binary=None, order=None):
    """One reference process per file, `cores` at a time, in `order` (longest first keeps every core busy
    to the end).  -> (GB/s, wall seconds, core-seconds, outputs|None)"""
    import pyoracle
    from concurrent.futures import ThreadPoolExecutor
    binary = binary or pyoracle.REF_BIN
    if not os.path.exists(binary):
        raise FileNotFoundError(binary)
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.mkdtemp(prefix="hcref_", dir=base)
    try:
        paths = []
        for i, f in enumerate(files):
            p = os.path.join(tmp, "f%05d.raw" % i)
            f.tofile(p)
            paths.append(p)
        idx = list(order) if order is not None else list(range(len(paths)))
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=cores) as ex:
            res = list(ex.map(_ref_one, [(binary, flags, paths[i]) for i in idx]))
        wall = time.perf_counter() - t0
        assert all(a == 0 and b == 0 for a, b, _ in res), "reference binary failed"
        outs = None
        if keep_outputs:
            outs = [np.fromfile(p + ".out", np.uint8) for p in paths]
            for p, f in zip(paths, files):
                assert np.array_equal(np.fromfile(p + ".dec", np.uint8), f), "reference round trip failed"
        nbytes = sum(f.size for f in files)
        return 2.0 * nbytes / wall / 1e9, wall, sum(t for _, _, t in res), outs
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# measured here (reference -O2, one core, compress + decompress, seconds per 512x512 file): used only to SIZE the sample
_CLASS_SECONDS = {"random": 1.5, "walk": 0.45, "smooth": 0.12, "const": 0.06}


def cpu_sample(wl, cores, budget_s):
    """Stratified sample of the workload (equal count per class) sized for ~budget_s seconds of wall time on
    `cores` cores, at least 4 files per core overall; returns (files, per-class count, start order)."""
    side, classes, seed0 = wl[0], wl[1], wl[2]
    scale = (side / 512.0) ** 2
    per_round = sum(_CLASS_SECONDS.get(c, 0.5) for c in classes) * scale       # core-seconds for one file of each class
    rounds = int(budget_s * cores / per_round)
    rounds = max(rounds, (4 * cores + len(classes) - 1) // len(classes) if side == 512 else 1)
    rounds = min(rounds, 4096 // len(classes))
    n = rounds * len(classes)
    files = [_gen_one((i, seed0, side, classes, False)) for i in range(n)]
    # longest-processing-time first: the slow classes start first, the fast ones fill the tail
    order = sorted(range(n), key=lambda i: -_CLASS_SECONDS.get(classes[i % len(classes)], 0.5))
    return files, rounds, order


def flags_of(wl):
    return ["-m"] + (["-a", "-w", str(wl[0])] if wl[3] else [])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    flags = flags_of(wl)
    nsteps = args.steps + min(args.warmup, 1)
    try:
        files, per_class, order = cpu_sample(wl, cores, min(25.0, 150.0 / max(nsteps, 1)))
        n = len(files)
        for _ in range(min(args.warmup, 1)):
            reference_sample(files, flags, cores, order=order)
        t_all = core_s = 0.0
        for _ in range(args.steps):
            _, wall, cs, _ = reference_sample(files, flags, cores, order=order)
            t_all += wall
            core_s += cs
        value = 2.0 * sum(f.size for f in files) * args.steps / t_all / 1e9
    except Exception as e:  # the oracle always exists; report why the reference arm could not run
        print(json.dumps({"impl": "reference", "unavailable": "%s: %s" % (type(e).__name__, e)}))
        return 0
    sample = ("%d files of the workload (%d per class, seeds %d+i), -c then -d, one process per file, %d at a time, slowest class first; "
              "core utilisation %.0f %%" % (n, per_class, wl[2], cores, 100.0 * core_s / (t_all * cores)))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": (wl[5] % wl[4]) + ", compress then decompress; bounded sample: " + sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
                         "core_seconds_per_step": core_s / args.steps},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                if t0 - 0.05 <= t <= t1 + 0.15:
                    sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            if t0 - 0.05 <= t <= t1 + 0.15:
                for nme, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ our arm
class DeviceArm:
    """One codec + its device-resident output buffers: compress / decompress of the whole local batch."""

    def __init__(self, torch, hc_b200, L, dev, d_in, d_in_off, d_in_len, d_width, nf, file_bytes, use_adapt):
        self.L, self.hc = L, hc_b200
        self.cd = hc_b200.Codec(dev.index, L)
        self.stream = torch.cuda.ExternalStream(L.hc_codec_stream(self.cd.h), device=dev)
        self.nf, self.fb, self.use_adapt = nf, file_bytes, use_adapt
        self.d_in, self.d_in_off, self.d_in_len, self.d_width = d_in, d_in_off, d_in_len, d_width
        i64 = torch.int64
        self.m_bound = file_bytes + file_bytes // 3 + 64 + (file_bytes // 8 + 512 if use_adapt else 0)
        self.cap = hc_b200.align_up(int(L.hc_fgk_bound(self.m_bound)) + 16)
        self.d_cmp = torch.empty(nf * self.cap + 512, dtype=torch.uint8, device=dev)
        self.d_cmp_off = torch.arange(nf, dtype=i64, device=dev) * self.cap
        self.d_cmp_cap = torch.full((nf,), self.cap, dtype=i64, device=dev)
        self.d_cmp_len = torch.zeros(nf, dtype=i64, device=dev)
        self.d_st_c = torch.zeros(nf, dtype=torch.int32, device=dev)
        self.d_dec = torch.empty(nf * file_bytes + 512, dtype=torch.uint8, device=dev)
        self.d_dec_len = torch.zeros(nf, dtype=i64, device=dev)
        self.d_st_d = torch.zeros(nf, dtype=torch.int32, device=dev)
        self.kinds = hc_b200.KIND_DIFF | (hc_b200.KIND_ADAPT if use_adapt else hc_b200.KIND_PLAIN)

    def compress(self):
        self.hc.check(self.L.hc_compress_device(self.cd.h, self.d_in.data_ptr(), self.d_in_off.data_ptr(), self.d_in_len.data_ptr(),
                                                self.d_width.data_ptr(), self.nf, self.fb, 1, int(self.use_adapt), self.d_cmp.data_ptr(),
                                                self.d_cmp_off.data_ptr(), self.d_cmp_cap.data_ptr(), self.d_cmp_len.data_ptr(),
                                                self.d_st_c.data_ptr()), "hc_compress_device", self.L)

    def decompress(self):
        self.hc.check(self.L.hc_decompress_device(self.cd.h, self.d_cmp.data_ptr(), self.d_cmp_off.data_ptr(), self.d_cmp_len.data_ptr(),
                                                  self.nf, self.m_bound, self.fb, self.kinds, self.d_dec.data_ptr(), self.d_in_off.data_ptr(),
                                                  self.d_in_len.data_ptr(), self.d_dec_len.data_ptr(), self.d_st_d.data_ptr()),
                      "hc_decompress_device", self.L)

    def stage_times(self):
        buf = (C.c_float * 16)()
        n = self.L.hc_codec_stage_times(self.cd.h, buf, 16)
        return {self.L.hc_stage_name(self.cd.h, i).decode(): float(buf[i]) for i in range(max(n, 0))}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import hc_b200
    import shard

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    wl = WORKLOADS[args.workload]
    side, use_adapt = wl[0], wl[3]
    FB = side * side
    strong = args.total_files is not None
    if strong:
        lo, hi = shard.shard_range(args.total_files, rank, world)
        total_files = args.total_files
    else:
        per = args.files if args.files is not None else wl[4]
        lo, hi = rank * per, (rank + 1) * per
        total_files = per * world
    nf = hi - lo
    # generate before CUDA init (fork pool)
    t_gen = time.perf_counter()
    host_batch, classes = make_batch(lo, hi, wl, (os.cpu_count() or 1) // max(1, world))
    t_gen = time.perf_counter() - t_gen

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        # NCCL announces its version on stdout at the first collective; keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
            comm = shard.NcclComm(rank, world)          # communicator for the C-ABI collective
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    L = hc_b200.lib()
    flags_txt = " ".join(flags_of(wl))

    # ---- device-resident buffers (kernel-only arm) ------------------------------------------
    i64 = torch.int64
    d_in = torch.from_numpy(host_batch).to(dev)                       # nf x FB, rows 256-aligned
    d_in_off = torch.arange(nf, dtype=i64, device=dev) * FB
    d_in_len = torch.full((nf,), FB, dtype=i64, device=dev)
    d_width = torch.full((nf,), side, dtype=i64, device=dev)
    n_arms = max(1, min(args.overlap, 8))
    # keep the resident buffers of all streams (and, later, of the e2e pipeline slots) well inside the 180 GB of HBM
    cap_est = FB + FB // 3 + FB // 8 + 8192
    per_stream = nf * (cap_est * 17 // 8 + 5 * FB)
    n_arms = max(1, min(n_arms, int(70e9 // max(per_stream, 1))))
    arms = [DeviceArm(torch, hc_b200, L, dev, d_in, d_in_off, d_in_len, d_width, nf, FB, use_adapt) for _ in range(n_arms)]
    A0 = arms[0]
    for arm in arms:                                                  # per stream: scratch + result of the size exchange
        arm.shard_ws = torch.empty(int(L.hc_shard_ws_bytes(total_files, world)) + 64, dtype=torch.uint8, device=dev)
        arm.g_sizes = torch.empty(total_files, dtype=i64, device=dev)
        arm.g_offs = torch.empty(total_files, dtype=i64, device=dev)
        arm.g_total = torch.zeros(1, dtype=i64, device=dev)

    def gather_sizes(arm):
        # the path's only collective: per-file output sizes -> global offsets table, through the C ABI.
        # Contiguous shards of the strong-scaling split differ by at most one file; the weak split is even.
        hc_b200.check(L.hc_shard_sizes_allgather(comm.comm if comm else None, rank, world, arm.d_cmp_len.data_ptr(), total_files, 16,
                                                 arm.g_sizes.data_ptr(), arm.g_offs.data_ptr(), arm.g_total.data_ptr(), arm.shard_ws.data_ptr(),
                                                 L.hc_codec_stream(arm.cd.h)), "hc_shard_sizes_allgather", L)

    def step(arm):
        arm.compress()
        gather_sizes(arm)
        arm.decompress()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up + parity check on every rank -------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        for arm in arms:
            step(arm)
    barrier()
    for arm in arms:
        assert int(arm.d_st_c.abs().sum().item()) == 0 and int(arm.d_st_d.abs().sum().item()) == 0, "non-zero status"
        assert torch.equal(arm.d_dec[: nf * FB].view(nf, FB), d_in), "round trip differs"
    assert torch.equal(arms[0].d_cmp_len, arms[-1].d_cmp_len)
    out_bytes = int(A0.d_cmp_len.sum().item())
    import pyoracle
    ora = pyoracle.Oracle()
    lens = A0.d_cmp_len.cpu().numpy()
    n_par = min(nf, 16 if side == 512 else 4)
    for i in range(n_par):                               # the first files of EVERY rank's shard against the oracle
        got = A0.d_cmp[i * A0.cap: i * A0.cap + int(lens[i])].cpu().numpy()
        rc, exp = ora.compress(host_batch[i], diff=True, adapt=use_adapt, width=side)
        assert rc == 0 and np.array_equal(got, exp), "GPU .out differs from the oracle (rank %d file %d)" % (rank, i)
    # the offsets table every rank derived equals the scan of the gathered sizes, and is the same on every rank
    g_sizes, g_offs, g_total = A0.g_sizes, A0.g_offs, A0.g_total
    chk = torch.cat([g_sizes.sum().view(1), g_offs[-1:].view(1), g_total.view(1)]).to(torch.float64)
    if world > 1:
        mx, mn = chk.clone(), chk.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        assert torch.equal(mx, mn), "ranks disagree on the offsets table"
    assert torch.equal(g_offs, torch.cumsum((g_sizes + 15) // 16 * 16, 0) - (g_sizes + 15) // 16 * 16)
    parity = torch.tensor([n_par], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(parity)
    parity_files = int(parity.item())

    # ---- timed region: device resident, consecutive steps alternate between the codec streams ----
    sampler = ClockSampler(local) if rank == 0 else None
    master = torch.cuda.current_stream(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = L.hc_launch_count()
    t0 = time.perf_counter()
    ev0.record(master)
    for arm in arms:
        arm.stream.wait_event(ev0)
    for s in range(args.steps):
        step(arms[s % n_arms])
    for arm in arms:
        e = torch.cuda.Event()
        e.record(arm.stream)
        master.wait_event(e)
    ev1.record(master)
    barrier()
    t1 = time.perf_counter()
    launches = L.hc_launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total = float(tt.item())
    clocks = sampler.stop(t0, t1) if sampler else None
    ms_step = ms_total / args.steps
    n_in = nf * FB
    total_in = total_files * FB
    value = 2.0 * total_in / (ms_step * 1e-3) / 1e9

    # ---- one stream, one step at a time: per-stage device times (the roofline numbers) -----------
    L.hc_codec_enable_stage_timing(A0.cd.h, 1)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    evs[0].record(A0.stream)
    seq_steps = max(1, min(args.steps, 3))
    for _ in range(seq_steps):
        step(A0)
    evs[1].record(A0.stream)
    barrier()
    seq_ms = evs[0].elapsed_time(evs[1]) / seq_steps
    st_dec = A0.stage_times()
    A0.compress()
    torch.cuda.synchronize()
    st_cmp = A0.stage_times()
    # the -m sub-run: plain MNP-5 RLE instead of the adaptive stage, so that rle_encode / rle_decode are measured too
    st_m = {}
    if use_adapt and side == 512 and not args.no_extras:
        Am = DeviceArm(torch, hc_b200, L, dev, d_in, d_in_off, d_in_len, d_width, nf, FB, False)
        L.hc_codec_enable_stage_timing(Am.cd.h, 1)
        for _ in range(2):
            Am.compress()
            Am.decompress()
        torch.cuda.synchronize()
        assert int(Am.d_st_c.abs().sum().item()) == 0 and torch.equal(Am.d_dec[: nf * FB].view(nf, FB), d_in)
        st_m = Am.stage_times()
        Am.compress()
        torch.cuda.synchronize()
        st_m.update(Am.stage_times())
        # symbols per file = first 8 bytes of each .out, ALL files (a strided sample would hit one class only)
        hdr_m = Am.d_cmp[: nf * Am.cap].view(nf, Am.cap)[:, :8].cpu().numpy()
        m_sym_plain = int((hdr_m.astype(np.uint64) @ (np.uint64(1) << (np.uint64(8) * np.arange(8, dtype=np.uint64)))).sum())
        del Am
        torch.cuda.empty_cache()
    L.hc_codec_enable_stage_timing(A0.cd.h, 0)

    # ---- saturated run: 4 x the batch per GPU (FGK issue- rather than latency-bound) -------------
    saturated = None
    if side == 512 and nf * 4 * FB <= 8 << 30 and not args.no_extras and not strong:
        rep = 4
        big_in = d_in.repeat(rep, 1)
        nfb = nf * rep
        Ab = DeviceArm(torch, hc_b200, L, dev, big_in, torch.arange(nfb, dtype=i64, device=dev) * FB, torch.full((nfb,), FB, dtype=i64, device=dev),
                       torch.full((nfb,), side, dtype=i64, device=dev), nfb, FB, use_adapt)
        L.hc_codec_enable_stage_timing(Ab.cd.h, 1)
        Ab.compress()
        Ab.decompress()
        eb = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        eb[0].record(Ab.stream)
        for _ in range(2):
            Ab.compress()
            Ab.decompress()
        eb[1].record(Ab.stream)
        torch.cuda.synchronize()
        assert int(Ab.d_st_c.abs().sum().item()) == 0 and torch.equal(Ab.d_dec[: nfb * FB].view(nfb, FB), big_in)
        msb = eb[0].elapsed_time(eb[1]) / 2
        sd = Ab.stage_times()
        Ab.compress()
        torch.cuda.synchronize()
        sc = Ab.stage_times()
        saturated = {"files_per_gpu": nfb, "what": "the batch replicated %d x on the device, one stream" % rep, "ms_per_step": msb,
                     "value": 2.0 * nfb * FB / (msb * 1e-3) / 1e9, "unit": UNIT, "fgk_encode_ms": sc.get("fgk_encode"), "fgk_decode_ms": sd.get("fgk_decode")}
        del Ab, big_in
        torch.cuda.empty_cache()

    # ---- e2e: host buffers through the asynchronous public API (depth-2 pipeline) --------------------
    e2e = None
    if not args.no_e2e:
        pin_in = torch.from_numpy(host_batch).pin_memory()
        offs = (np.arange(nf, dtype=np.uint64) * FB)
        lens_h = np.full(nf, FB, np.uint64)
        widths = np.full(nf, side, np.uint64)
        pipe = C.c_void_p()
        depth = max(2, min(args.depth, 8, int(60e9 // max(per_stream, 1)))) & ~1
        nbuf = depth + 1
        pin_cmp = [torch.empty(out_bytes + 16 * nf + 4096, dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
        pin_dec = [torch.empty(nf * FB + 4096, dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
        tabs = [[np.zeros(nf, np.uint64), np.zeros(nf, np.uint64), np.zeros(nf, np.int32), np.zeros(nf, np.uint64), np.zeros(nf, np.uint64),
                 np.zeros(nf, np.int32)] for _ in range(nbuf)]
        hc_b200.check(L.hc_pipeline_create(C.byref(pipe), local, depth), "hc_pipeline_create", L)

        def sub_c(k):
            o_off, o_len, o_st = tabs[k][:3]
            t = L.hc_pipeline_submit_compress(pipe, pin_in.data_ptr(), offs.ctypes.data, lens_h.ctypes.data, nf, 1, int(use_adapt),
                                              widths.ctypes.data, pin_cmp[k].data_ptr(), pin_cmp[k].numel(), o_off.ctypes.data, o_len.ctypes.data,
                                              o_st.ctypes.data)
            assert t >= 0
            return t

        def sub_d(k):
            o_off, o_len, _, r_off, r_len, r_st = tabs[k]
            t = L.hc_pipeline_submit_decompress(pipe, pin_cmp[k].data_ptr(), o_off.ctypes.data, o_len.ctypes.data, nf, pin_dec[k].data_ptr(),
                                                pin_dec[k].numel(), r_off.ctypes.data, r_len.ctypes.data, r_st.ctypes.data)
            assert t >= 0
            return t

        def run_host(k_steps):
            # step s: compress, then decompress of its output.  Tickets alternate compress / decompress so that the two
            # kinds land on different pipeline slots; up to depth/2 steps are in flight, a host buffer set is reused only
            # after the decompress that read it has finished.
            ahead = max(1, min(args.ahead or depth // 2, depth // 2))
            tc, td = {}, {}
            nxt = 0                                           # next step whose compress is to be submitted
            for s in range(k_steps):
                while nxt < k_steps and nxt < s + ahead:
                    if nxt - nbuf in td:
                        hc_b200.check(L.hc_pipeline_wait(pipe, td.pop(nxt - nbuf)), "decompress job", L)
                    tc[nxt] = sub_c(nxt % nbuf)
                    nxt += 1
                hc_b200.check(L.hc_pipeline_wait(pipe, tc.pop(s)), "compress job", L)
                td[s] = sub_d(s % nbuf)
            for s in sorted(td):
                hc_b200.check(L.hc_pipeline_wait(pipe, td[s]), "decompress job", L)

        run_host(nbuf + 1)                                        # warm-up: the buffers of every slot allocated, every host buffer set used
        for k in range(nbuf):
            assert not tabs[k][2].any() and not tabs[k][5].any()
            assert np.array_equal(pin_dec[k].numpy()[: nf * FB].reshape(nf, FB), host_batch), "e2e round trip differs"
        barrier()
        th0 = time.perf_counter()
        run_host(args.steps)
        barrier()
        th = (time.perf_counter() - th0) / args.steps
        tth = torch.tensor([th], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tth, op=dist.ReduceOp.MAX)
        th = float(tth.item())
        comp_bytes = int(tabs[0][1].sum())
        L.hc_pipeline_destroy(pipe)
        e2e = {"value": 2.0 * total_in / th / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(nf * FB + comp_bytes), "d2h_bytes_per_step": int(comp_bytes + nf * FB), "steps": args.steps,
               "ms_per_step": th * 1e3, "timer": "host wall clock around %d steps through hc_pipeline_submit_compress/_decompress + hc_pipeline_wait "
               "(depth %d: the copies and kernels of neighbouring steps overlap), max over ranks" % (args.steps, depth)}

    if rank != 0:
        if comm:
            comm.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline ---------------------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # symbols fed to FGK per file = first 8 bytes of each .out
    hdr = A0.d_cmp[: nf * A0.cap].view(nf, A0.cap)[:, :8].cpu().numpy()
    m_per_file = hdr.astype(np.uint64) @ (np.uint64(1) << (np.uint64(8) * np.arange(8, dtype=np.uint64)))
    m_sym = int(m_per_file.sum())
    alg = {  # algorithmic bytes per launch: what the stage must read + write once
        "diff_apply": 2 * n_in, "diff_revert": 2 * n_in, "adapt_encode": n_in + m_sym, "adapt_decode": m_sym + n_in,
        "rle_encode": n_in + m_sym, "rle_decode": m_sym + n_in, "fgk_encode": m_sym + out_bytes, "fgk_decode": out_bytes + m_sym,
    }
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath) and args.workload == "c3ma" and nf == 4096:   # per-launch DRAM bytes from the committed ncu capture
        traffic = json.load(open(tpath))
    stages = []
    for name, ms in list(st_cmp.items()) + list(st_dec.items()):
        if name in alg and ms > 0:
            a = alg[name] / (ms * 1e-3) / 1e9
            stages.append({"kernel": name, "bound": "hbm", "ms": ms, "algorithmic_bytes": alg[name], "achieved": a, "peak": peak, "unit": "GB/s",
                           "frac": a / peak, "traffic": traffic.get(name)})
    for name in ("rle_encode", "rle_decode"):
        if st_m.get(name, 0) > 0:
            ab = n_in + m_sym_plain
            a = ab / (st_m[name] * 1e-3) / 1e9
            stages.append({"kernel": name, "bound": "hbm", "ms": st_m[name], "algorithmic_bytes": ab, "achieved": a, "peak": peak, "unit": "GB/s",
                           "frac": a / peak, "traffic": traffic.get(name), "from": "-m sub-run of the same batch (plain MNP-5 RLE instead of -a)"})
    # FGK: bound by instruction issue / the dependent chain of one warp, not by HBM (north_star).  Warp instructions
    # per symbol per class come from the committed ncu capture of THESE kernels (profiles/r02_fgk_inst_per_symbol.json);
    # the instruction count of this run is derived from them and this run's symbol counts, the time is measured here.
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    issue_peak = 148 * 4 * sm_mhz * 1e6 / 1e9                       # G warp-instructions / s: one issue slot per scheduler and cycle
    ips = None
    ipath = os.path.join(ROOT, "profiles", "r02_fgk_inst_per_symbol.json")
    if os.path.exists(ipath):
        ips = json.load(open(ipath))
    cls_of = [classes[(lo + i) % len(classes)] for i in range(nf)]
    fgk = {}
    for nm, d in (("fgk_encode", st_cmp), ("fgk_decode", st_dec)):
        if nm in d and d[nm] > 0:
            ent = {"ms": d[nm], "symbols": m_sym, "streams": nf, "symbols_per_s": m_sym / (d[nm] * 1e-3), "streams_per_s": nf / (d[nm] * 1e-3),
                   "ns_per_symbol_longest_stream": d[nm] * 1e6 / float(m_per_file.max())}
            if ips and nm in ips and all(c in ips[nm] for c in set(cls_of)):
                winst = float(sum(float(m_per_file[i]) * ips[nm][cls_of[i]] for i in range(nf)))
                ach = winst / (d[nm] * 1e-3) / 1e9
                ent.update({"bound": "issue", "warp_instructions": winst, "achieved": ach, "peak": issue_peak, "unit": "G warp-inst/s", "frac": ach / issue_peak,
                            "inst_per_symbol": ips[nm], "inst_source": "derived: ncu smsp__inst_executed.sum per symbol and class (profiles/r02_fgk_inst_per_symbol.json) x the symbols of this run"})
            fgk[nm] = ent
    dom_name = max(fgk, key=lambda k: fgk[k]["ms"]) if fgk else None
    roofline = None
    if dom_name and "frac" in fgk[dom_name]:
        dm = fgk[dom_name]
        roofline = {"kernel": dom_name, "bound": "issue", "achieved": dm["achieved"], "peak": dm["peak"], "unit": dm["unit"], "frac": dm["frac"],
                    "traffic": traffic.get(dom_name), "peak_source": "148 SMs x 4 schedulers x %.0f MHz (SM clock sampled during the timed region)" % sm_mhz,
                    "hbm": {"achieved": alg[dom_name] / (dm["ms"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "peak_source": peak_src},
                    "note": "FGK is serial per stream: a batch is bound by instruction issue and by the dependent chain of the longest stream's warp, "
                            "so its roofline is the issue rate (north_star); the HBM-bound transform kernels are listed in `stages`",
                    "stages": stages}
    elif stages:
        dom = max(stages, key=lambda s: s["ms"])
        roofline = dict(dom, stages=stages, peak_source=peak_src)

    # ---- CPU baseline (bounded sample, rank 0, N == 1 only) -----------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        try:
            files, per_class, order = cpu_sample(wl, cores, 20.0)
            n = len(files)
            v, wall, core_s, outs = reference_sample(files, flags_of(wl), cores, keep_outputs=True, order=order)
            # the same files through the GPU arm must give the reference's bytes
            if not _CLASSES_ENV and not (wl is WORKLOADS["c5"]):
                lens_c = A0.d_cmp_len.cpu().numpy()
                for i in range(min(n, nf)):
                    got = A0.d_cmp[i * A0.cap: i * A0.cap + int(lens_c[i])].cpu().numpy()
                    assert np.array_equal(got, outs[i]), "GPU .out differs from the reference binary (file %d)" % i
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "core_seconds": core_s, "core_utilisation": core_s / (wall * cores),
                   "sample": "%d of the %d files (%d per class), reference binary -O2, -c then -d, one process per file, %d at a time, slowest class "
                             "first, %.1f s wall; GPU outputs byte-identical on the sample" % (n, nf, per_class, cores, wall)}
        except FileNotFoundError:
            files = [_gen_one((i, wl[2], side, classes, False)) for i in range(8)]
            tc = time.perf_counter()
            for f in files:
                rc, o = ora.compress(f, diff=True, adapt=use_adapt, width=side, mode=0)
                ora.decompress(o, mode=0)
            wall = time.perf_counter() - tc
            cpu = {"value": 2.0 * sum(f.size for f in files) / wall / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "8 files (2 per class) through oracle/hc_oracle.c mode 0 (faithful tree), single thread, %.1f s" % wall}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": (wl[5] % nf) + ", compress then decompress", "flags": flags_txt,
                   "files_per_gpu": nf, "total_files": total_files, "bytes_per_gpu": n_in, "overlap": n_arms,
                   "l2": "inputs (%.2f GiB per GPU) larger than L2, no flush needed" % (n_in / 2 ** 30),
                   "parallelism": "files sharded over %d GPU(s), contiguous shards; hc_shard_sizes_allgather (NCCL) of the per-file sizes" % world},
        "sequential_ms_per_step": seq_ms, "sequential_value": 2.0 * total_in / (seq_ms * 1e-3) / 1e9,
        "bpc": 8.0 * out_bytes / n_in, "compressed_bytes_rank0": out_bytes, "parity_checked_files": parity_files,
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "fgk": fgk, "saturated": saturated,
        "cpu_baseline": cpu, "stage_ms": {"compress": st_cmp, "decompress": st_dec, "m_subrun": st_m}, "gen_seconds": t_gen,
    }
    print(json.dumps(line))
    if comm:
        comm.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3ma", choices=sorted(WORKLOADS))
    ap.add_argument("--files", type=int, default=None, help="files per GPU (weak scaling; default: the workload's batch)")
    ap.add_argument("--total-files", type=int, default=None, help="strong scaling: this many files in total, cut over the ranks")
    ap.add_argument("--overlap", type=int, default=8, help="codec streams that consecutive steps alternate between (1 = strictly one step at a time, max 8)")
    ap.add_argument("--depth", type=int, default=4, help="slots of the asynchronous host pipeline used by the e2e measurement (even: compress and decompress jobs alternate)")
    ap.add_argument("--ahead", type=int, default=0, help="steps in flight in the e2e measurement (default and maximum: depth / 2)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the -m sub-run and the saturated run")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
