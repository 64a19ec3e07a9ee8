#!/usr/bin/env python
"""Benchmark of the hot path: batched VE posterior queries (headline) + CPT-fit counting.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference-style CPU path (oracle port)

Headline workload = BASELINE.json configs[1]: Asia (8 binary nodes), 1,048,576 evidence rows per GPU,
evidence {asia, smoke, xray, dysp} drawn from the joint, targets lung / tub / bronc.  One *step* = the three
compiled plans run over one batch of evidence rows; one *query* = one posterior row of one target.
Inputs are device resident for `value`; `e2e` runs the same step through the C-ABI host-buffer call
(pinned host codes in, pinned host posteriors out, copies inside the timed region).
Between timed iterations the step rotates through a ring of distinct batches larger than L2; the K timed steps are
ONE CUDA graph launch (kernels of consecutive steps spread over 3 streams inside the graph), `roofline.one_stream`
reports the same launches as one graph on a single stream.
Extra keys carry the other configurations: `fit` / `fit_e2e` / `ingest_fit_f32` (CPT counting), `alarm`, `alarm_fit`,
`ktree200_fit`, `ktree200_ve`, `layered1000_ve` (BASELINE.json configs 3-5).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ASIA_EVIDENCE = ["asia", "smoke", "xray", "dysp"]
ASIA_TARGETS = ["lung", "tub", "bronc"]
ROWS_PER_GPU = 1 << 20
L2_BYTES = 126 * 1024 * 1024
METRIC = "posterior queries/sec (batched VE)"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _profile_traffic(kernel_key):
    """dram bytes per launch of the dominant kernel from the committed ncu summary (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(kernel_key)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            parts = [x.strip() for x in l.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- reference arm
def cpu_ve_queries_per_sec(budget_s, rows_per_call=65536, seed=11):
    """The reference-style CPU path for a multi-layer network (oracle O2: textbook batched VE in
    fp32 PyTorch on the host cores), on a bounded sample of the Asia workload.  Returns (queries/s, sample text, cores)."""
    import numpy as np
    import torch

    from continuousbayesiannetwork_b200 import synth
    from oracle import cbn_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    spec = synth.asia()
    n_fit = 200_000
    codes = synth.sample_forward_numpy(spec, 1235, 0, n_fit)
    fitted = [O.cpt_from_counts(O.dense_counts(codes, spec.parents[i] + [i], spec.cards), n_fit)[1] for i in range(spec.n)]
    net = O.DiscreteNet(spec.cards, spec.parents, fitted)
    ids = [spec.names.index(e) for e in ASIA_EVIDENCE]
    ev = synth.sample_forward_numpy(spec, seed, 0, rows_per_call)[ids].T
    O.ve_posterior(net, spec.names.index("lung"), ids, ev[:1024])           # warm-up
    done, t0 = 0, time.perf_counter()
    while True:
        for t in ASIA_TARGETS:
            O.ve_posterior(net, spec.names.index(t), ids, ev)
        done += rows_per_call * len(ASIA_TARGETS)
        el = time.perf_counter() - t0
        if el >= budget_s:
            break
    return done / el, f"{done} queries = {done // (rows_per_call * 3)} x ({rows_per_call} rows x 3 targets) of the Asia workload in {el:.1f}s", torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = 262144
    import numpy as np
    import torch

    from continuousbayesiannetwork_b200 import synth
    from oracle import cbn_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    spec = synth.asia()
    n_fit = 200_000
    codes = synth.sample_forward_numpy(spec, 1235, 0, n_fit)
    fitted = [O.cpt_from_counts(O.dense_counts(codes, spec.parents[i] + [i], spec.cards), n_fit)[1] for i in range(spec.n)]
    net = O.DiscreteNet(spec.cards, spec.parents, fitted)
    ids = [spec.names.index(e) for e in ASIA_EVIDENCE]
    ev = synth.sample_forward_numpy(spec, 11, 0, rows)[ids].T

    def step():
        for t in ASIA_TARGETS:
            O.ve_posterior(net, spec.names.index(t), ids, ev)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    q = rows * len(ASIA_TARGETS) * args.steps
    val = q / el
    sample = f"each step = {rows} evidence rows x 3 targets of the Asia workload (bounded sample of the 1,048,576-row batch)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "asia-8node, evidence {asia,smoke,xray,dysp}, targets lung/tub/bronc (BASELINE.json configs[1])",
                   "rows_per_step": rows, "path": "reference-style PyTorch CPU path: oracle O2 textbook batched VE fp32 "
                   "(the reference has no correct multi-layer query path, SURVEY.md 3.3)"},
        "cpu_baseline": {"value": val, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def _bind_to_gpu_numa_node(index):
    """One process per GPU: run (and first-touch the pinned staging memory) on the CPU cores next to that GPU, so the
    host<->device copies of the e2e path do not cross the socket interconnect."""
    before = None
    try:
        import pynvml

        before = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [w * 64 + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1 and w * 64 + b < n_cpu]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
    except Exception:
        pass
    return before


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from continuousbayesiannetwork_b200 import sharding, synth
    from continuousbayesiannetwork_b200.engine import bind_inference, sample_network, tables_from_spec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    all_cpus = _bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak_gbs, peak_src = _peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup):
        """W untimed steps, then exactly K steps bracketed by barrier + synchronize; CUDA events on the launch
        stream; MAX over ranks.  Returns seconds."""
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        return max_over_ranks(ms) / 1e3

    # ---------------- fit (config 2 shape): 1e7 forward samples sharded over the ranks, one int64 all-reduce
    spec = synth.asia()
    n_fit = 10_000_000
    s, e = sharding.shard_range(n_fit, rank, world)
    tables = tables_from_spec(spec, dev)
    fit_codes = sample_network(spec, seed=1235, first=s, n=e - s, device=dev, tables=tables)
    torch.cuda.synchronize()

    # throughput is measured on a larger resident block (2^28 samples per GPU, 2 GB of codes >> L2) ...
    n_big = args.fit_samples
    big_tables = tables_from_spec(spec, dev)
    big_codes = sample_network(spec, seed=1235, first=n_fit + rank * n_big, n=n_big, device=dev, tables=big_tables)
    torch.cuda.synchronize()

    def fit_step(_i):
        big_tables.reset_counts()
        sharding.fit_sharded(big_tables, big_codes, n_big)

    fit_steps = max(3, min(args.steps, 20))
    fit_s = timed(fit_step, fit_steps, 3)
    fit_rate = n_big * world * fit_steps / fit_s
    fit_updates = big_tables.count_updates_per_sample()
    del big_codes, big_tables
    # ... the CPTs the queries use come from the configuration's 1e7 samples, sharded over the ranks
    sharding.fit_sharded(tables, fit_codes, e - s)
    infer = bind_inference(tables)

    # ---------------- queries: ring of distinct batches larger than L2
    rows = args.rows
    ld = (rows + 15) // 16 * 16
    bytes_per_batch = len(ASIA_EVIDENCE) * ld + len(ASIA_TARGETS) * rows * 2 * 4
    ring = max(2, -(-2 * L2_BYTES // bytes_per_batch))
    ids = [spec.names.index(x) for x in ASIA_EVIDENCE]
    ev_ring, out_ring = [], []
    for r in range(ring):
        # evidence rows drawn from the joint: forward samples with a per-rank, per-batch counter range
        full = sample_network(spec, seed=4321, first=(rank * ring + r) * rows, n=rows, device=dev, tables=tables)
        ev_ring.append(full[ids].contiguous())
        out_ring.append([torch.empty((rows, 2), dtype=torch.float32, device=dev) for _ in ASIA_TARGETS])
    del full
    t_compile = time.perf_counter()
    plans = [infer.plan(t, ASIA_EVIDENCE) for t in ASIA_TARGETS]
    fused = infer.fused_plan(ASIA_TARGETS, ASIA_EVIDENCE)      # one launch answers the three targets
    torch.cuda.synchronize()
    compile_ms = (time.perf_counter() - t_compile) * 1e3
    for r in range(ring):                                       # warm-up outside any capture
        fused.run_codes(ev_ring[r], rows, outs=out_ring[r])
    torch.cuda.synchronize()
    # the step is launch-latency scale (tens of MB): replay it from CUDA graphs, one per ring slot plus one
    # holding a whole ring cycle, so the GPU is not waiting on Python between 10-microsecond kernels
    slot_graphs = []
    for r in range(ring):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fused.run_codes(ev_ring[r], rows, outs=out_ring[r])
        slot_graphs.append(g)
    # whole ring cycle: the steps are independent batches, so inside the graph they are spread over three
    # streams (what a serving loop does) and the launch latency of one step hides behind the previous one
    cycle_graph = torch.cuda.CUDAGraph()
    side = [torch.cuda.Stream(device=dev) for _ in range(args.graph_streams - 1)]
    with torch.cuda.graph(cycle_graph):
        main = torch.cuda.current_stream()
        for st_ in side:
            st_.wait_stream(main)
        for r in range(ring):
            k = r % args.graph_streams
            if k == 0:
                fused.run_codes(ev_ring[r], rows, outs=out_ring[r])
            else:
                with torch.cuda.stream(side[k - 1]):
                    fused.run_codes(ev_ring[r], rows, outs=out_ring[r])
        for st_ in side:
            main.wait_stream(st_)

    def run_query_steps(first, count):
        """`count` consecutive steps starting at ring position first % ring."""
        i, end = first, first + count
        while i < end and i % ring:
            slot_graphs[i % ring].replay(); i += 1
        while end - i >= ring:
            cycle_graph.replay(); i += ring
        while i < end:
            slot_graphs[i % ring].replay(); i += 1

    def capture_steps(first, count):
        """ONE graph holding exactly the steps first .. first+count-1 (ring order), spread over the streams like the
        ring-cycle graph: a single graph launch, so a short run is not dominated by per-graph launch latency."""
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            main = torch.cuda.current_stream()
            for st_ in side:
                st_.wait_stream(main)
            for i in range(first, first + count):
                r, k = i % ring, (i - first) % args.graph_streams
                if k == 0:
                    fused.run_codes(ev_ring[r], rows, outs=out_ring[r])
                else:
                    with torch.cuda.stream(side[k - 1]):
                        fused.run_codes(ev_ring[r], rows, outs=out_ring[r])
            for st_ in side:
                main.wait_stream(st_)
        return g

    def timed_queries(steps, warmup):
        run_query_steps(0, warmup)
        timed_graph = capture_steps(warmup, steps) if steps <= 4096 else None
        if timed_graph is not None:
            timed_graph.replay()            # untimed: uploads the graph (extra warm-up beyond the W steps above)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if timed_graph is not None:
            timed_graph.replay()
        else:
            run_query_steps(warmup, steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        return max_over_ranks(ms) / 1e3

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    q_s = timed_queries(args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    queries_per_step = rows * len(ASIA_TARGETS) * world
    value = queries_per_step * args.steps / q_s
    launches = args.steps * world
    # the same K steps as ONE graph on ONE stream: consecutive launches only overlap through programmatic dependent
    # launch (prologue and evidence loads of launch i+1 behind the tail of launch i); reported beside the 3-stream figure
    if args.steps <= 4096:
        one = torch.cuda.CUDAGraph()
        with torch.cuda.graph(one):
            for i in range(args.warmup, args.warmup + args.steps):
                fused.run_codes(ev_ring[i % ring], rows, outs=out_ring[i % ring])
        serial = one.replay
    else:
        def serial():
            for i in range(args.warmup, args.warmup + args.steps):
                slot_graphs[i % ring].replay()
    serial()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    serial()
    e1.record()
    torch.cuda.synchronize()
    serial_s = max_over_ranks(e0.elapsed_time(e1)) / 1e3
    barrier()
    # roofline of the dominant kernel (fused gather): algorithmic bytes = relevant evidence codes in (once) +
    # one fp32 posterior per target out
    alg_bytes = rows * fused.algorithmic_bytes_per_row()
    launch_s = q_s / args.steps
    achieved = alg_bytes / launch_s / 1e9

    # ---------------- e2e through the C-ABI host-buffer call (pinned host memory in and out)
    host_ev = [x.cpu().pin_memory() for x in ev_ring[:2]]
    host_out = [[torch.empty((rows, 2), dtype=torch.float32).pin_memory() for _ in ASIA_TARGETS] for _ in range(2)]

    def e2e_step(i):
        r = i % 2
        fused.run_codes_host(host_ev[r], rows, host_out[r])

    e2e_steps = max(3, min(args.steps, 50))
    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = queries_per_step * e2e_steps / e2e_s
    # parity spot check of what the e2e call returned against the device path
    assert torch.equal(host_out[0][0], out_ring[0][0].cpu()) or args.steps < 1

    # ---------------- e2e through the reference-facing Python API: infer(target, {name: float tensor [nq,1]})
    f_ev = {nm: host_ev[0][k, :rows].to(torch.float32).reshape(-1, 1).pin_memory() for k, nm in enumerate(ASIA_EVIDENCE)}

    api_out = [torch.empty((rows, 2), dtype=torch.float32).pin_memory() for _ in ASIA_TARGETS]

    def api_step(_i):
        for tname, o in zip(ASIA_TARGETS, api_out):
            o.copy_(infer.infer(tname, f_ev), non_blocking=True)
        torch.cuda.synchronize()

    api_step(0)
    barrier()
    t0 = time.perf_counter()
    api_steps = 5
    for i in range(api_steps):
        api_step(i)
    torch.cuda.synchronize()
    api_s = max_over_ranks(time.perf_counter() - t0)
    api_value = queries_per_step * api_steps / api_s

    # the same three targets through the multi-target call: evidence uploaded and encoded once, one fused launch
    def many_step(_i):
        res = infer.infer_many(ASIA_TARGETS, f_ev)
        for tname, o in zip(ASIA_TARGETS, api_out):
            o.copy_(res[tname], non_blocking=True)
        torch.cuda.synchronize()

    many_step(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(api_steps):
        many_step(i)
    torch.cuda.synchronize()
    many_s = max_over_ranks(time.perf_counter() - t0)
    many_value = queries_per_step * api_steps / many_s
    assert torch.equal(api_out[0], host_out[0][0])

    extras = {}
    if not args.no_extras:
        extras = run_extras(args, dev, rank, world, timed, peak_gbs)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            if all_cpus:
                os.sched_setaffinity(0, all_cpus)      # the CPU baseline gets every host core
            v, sample, cores = cpu_ve_queries_per_sec(args.cpu_budget_s)
            cpu = {"value": v, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": q_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "asia-8node batched VE, evidence {asia,smoke,xray,dysp}, targets lung/tub/bronc "
                                   "(BASELINE.json configs[1])",
                       "rows_per_gpu_per_step": rows, "queries_per_step": queries_per_step,
                       "cpts": f"fitted on the GPU from {n_fit} forward samples (count kernel + int64 all-reduce)",
                       "l2": f"ring of {ring} distinct batches, {ring * bytes_per_batch / 1e6:.0f} MB > 2 x 126 MB L2",
                       "launch": f"1 fused kernel per step; the K timed steps are ONE CUDA graph launch ({args.graph_streams} streams inside the graph)", "plan_compile_ms": compile_ms},
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": len(ASIA_EVIDENCE) * rows * world,
                    "d2h_bytes_per_step": len(ASIA_TARGETS) * rows * 2 * 4 * world,
                    "call": "cbn_ve_run_codes_host_multi (pinned host uint8 codes in, 3 pinned host fp32 posteriors out)",
                    "timing": "host clock around the synchronous calls (each returns after its last D2H copy has landed), barrier + synchronize on both sides, max over ranks",
                    "python_api": {"value": api_value, "unit": "queries/s",
                                   "call": "ExactInference.infer(target, {name: pinned float32 [nq,1]}) -> pinned host copy"},
                    "python_api_many": {"value": many_value, "unit": "queries/s",
                                        "call": "ExactInference.infer_many(targets, {name: pinned float32 [nq,1]}) -> pinned host copies"}},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                         "traffic": _profile_traffic("gather_inter_kernel<2,3>"), "kernel": "gather_inter_kernel<2,3> (3 binary targets fused, interleaved table)",
                         "algorithmic_bytes_per_launch": alg_bytes, "launch_us": launch_s * 1e6, "peak_source": peak_src,
                         "frac_excluding_l2_tail": max(alg_bytes * args.steps - L2_BYTES, 0) / q_s / 1e9 / peak_gbs,
                         "note": "peak is the copy-measured HBM figure; this kernel's traffic is 86 % writes, and up to one L2 (126 MB) of its "
                                 "posteriors can still be dirty in L2 when the timed region ends: frac_excluding_l2_tail discounts those bytes",
                         "one_stream": {"launch_us": serial_s / args.steps * 1e6, "achieved": alg_bytes / (serial_s / args.steps) / 1e9,
                                        "frac": alg_bytes / (serial_s / args.steps) / 1e9 / peak_gbs,
                                        "note": "the same K launches as one graph on ONE stream (overlap only through programmatic dependent launch)"}},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "fit": {"metric": "CPT-fit samples/sec", "value": fit_rate, "unit": "samples/s", "n_samples_per_gpu_per_step": n_big,
                    "n_vars": spec.n, "table_updates_per_sample": fit_updates, "achieved_GBs": fit_rate * spec.n / 1e9, "frac_of_hbm_peak": fit_rate * spec.n / 1e9 / (peak_gbs * world),
                    "includes": "count kernel + int64 all-reduce + CPT normalisation"},
        }
        line.update(extras)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_extras(args, dev, rank, world, timed, peak_gbs):
    """Other BASELINE.json configurations, reported beside the headline (not the bench value)."""
    import torch

    from continuousbayesiannetwork_b200 import sharding, synth
    from continuousbayesiannetwork_b200.engine import bind_inference, install_cpts, sample_network, tables_from_spec

    out = {}

    def graphed(step_fn):
        """Capture one step (kernel launches only) in a CUDA graph: the extras time the GPU work, not the Python and
        ctypes overhead of issuing tens of small launches."""
        step_fn(0)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step_fn(0)
        return lambda _i: g.replay()

    # ingestion: float32 columns resident on the device -> domains -> codes -> counts -> CPTs, the path
    # behind BayesianNetwork(dag, data) once the frame is on the GPU (Asia shape, 2^24 rows x 8 columns)
    from continuousbayesiannetwork_b200.tables import DiscreteTables

    spec = synth.asia()
    n_in = 1 << 24
    t0_ = tables_from_spec(spec, dev)
    c_ = sample_network(spec, seed=99, first=rank * n_in, n=n_in, device=dev, tables=t0_)
    cols = {nm: (c_[i, :n_in].to(torch.float32) * 0.5 - 1.0) for i, nm in enumerate(spec.names)}
    del c_, t0_
    ing = DiscreteTables(spec.names, spec.parents_by_name(), device=dev)

    def istep(_i):
        ing.fit_columns(cols)

    sec = timed(istep, 5, 2)
    out["ingest_fit_f32"] = {"metric": "CPT-fit samples/sec from float32 columns (domain discovery + encoding + counting + CPTs)",
                             "value": n_in * world * 5 / sec, "unit": "samples/s", "rows_per_gpu": n_in, "n_vars": spec.n,
                             "bytes_per_value": "4 (domain scan) + 4 + 1 (encode) + 1 (count)",
                             "achieved_GBs": n_in * spec.n * 10 * world * 5 / sec / 1e9,
                             "frac_of_hbm_peak": n_in * spec.n * 10 * 5 / sec / 1e9 / peak_gbs}
    del cols, ing
    torch.cuda.empty_cache()
    # fit end to end from HOST codes (pinned): chunked H2D overlapped with counting, then CPTs (Asia, 2^26 samples)
    spec = synth.asia()
    n_h = 1 << 26
    th = tables_from_spec(spec, dev)
    hcodes = sample_network(spec, seed=98, first=rank * n_h, n=n_h, device=dev, tables=th).cpu().pin_memory()
    th.count_host(hcodes, n_h)
    torch.cuda.synchronize()
    barrier_t0 = time.perf_counter()
    for _ in range(3):
        th.reset_counts()
        th.count_host(hcodes, n_h)
        th.finalize()
    torch.cuda.synchronize()
    el = time.perf_counter() - barrier_t0
    out["fit_e2e"] = {"metric": "CPT-fit samples/sec, pinned host codes in (H2D inside the timed region)", "value": n_h * 3 / el * world,
                      "unit": "samples/s", "samples_per_gpu": n_h, "n_vars": spec.n, "h2d_bytes_per_step": n_h * spec.n,
                      "pcie_GBs": n_h * spec.n * 3 / el / 1e9, "call": "cbn_count_run_host + cbn_cpt_from_plan_dev"}
    del hcodes, th
    # config 3: Alarm-shaped, 16M evidence rows sharded over the ranks (strong scaling inside this extra)
    spec = synth.alarm()
    tables, infer = install_cpts(spec, dev)
    total_rows = 1 << 24
    s, e = sharding.shard_range(total_rows, rank, world)
    rows = e - s
    ids = [spec.names.index(x) for x in synth.ALARM_EVIDENCE]
    full = sample_network(spec, seed=777, first=s, n=rows, device=dev, tables=tables)
    ev = full[ids].contiguous()
    del full
    t0 = time.perf_counter()
    plans = [infer.plan(t, synth.ALARM_EVIDENCE) for t in synth.ALARM_TARGETS]
    torch.cuda.synchronize()
    compile_ms = (time.perf_counter() - t0) * 1e3
    outs = [torch.empty((rows, p.card_t), dtype=torch.float32, device=dev) for p in plans]
    fused = infer.fused_plan(synth.ALARM_TARGETS, synth.ALARM_EVIDENCE)

    def step(_i):
        fused.run_codes(ev, rows, outs=outs)

    k = max(3, min(args.steps, 10))
    sec = timed(graphed(step), k, 3)
    alg = rows * fused.algorithmic_bytes_per_row() * world
    out["alarm"] = {"metric": METRIC, "value": total_rows * len(plans) * k / sec, "unit": "queries/s",
                    "rows_total": total_rows, "targets": len(plans), "achieved_GBs": alg * k / sec / 1e9,
                    "frac_of_hbm_peak": alg * k / sec / 1e9 / (peak_gbs * world), "plan_compile_ms": compile_ms,
                    "table_cells": [p.stats.final_tables[0][1] for p in plans]}
    # fused MAP prediction (benchmarking_df path): one float per row instead of a posterior row
    mplan = plans[0]
    mout = torch.empty(rows, dtype=torch.float32, device=dev)

    def mstep(_i):
        mplan.run_codes_map(ev, rows, out=mout)

    sec = timed(graphed(mstep), k, 3)
    mbytes = rows * (len(mplan.stats.relevant_evidence) + 4) * world
    out["alarm_map"] = {"metric": "MAP predictions/sec (fused posterior + argmax + domain lookup)", "value": total_rows * k / sec,
                        "unit": "rows/s", "rows_total": total_rows, "target": synth.ALARM_TARGETS[0],
                        "achieved_GBs": mbytes * k / sec / 1e9, "frac_of_hbm_peak": mbytes * k / sec / 1e9 / (peak_gbs * world)}
    # config 3, fit half: counting on the Alarm structure
    n_chunk = 2 * args.fit_chunk
    codes = sample_network(spec, seed=1236, first=rank * n_chunk, n=n_chunk, device=dev, tables=tables)
    ftab = tables_from_spec(spec, dev)
    torch.cuda.synchronize()

    def astep(_i):
        ftab.reset_counts()
        sharding.fit_sharded(ftab, codes, n_chunk)

    k = max(3, min(args.steps, 10))
    sec = timed(astep, k, 2)
    rate = n_chunk * world * k / sec
    out["alarm_fit"] = {"metric": "CPT-fit samples/sec", "value": rate, "unit": "samples/s", "n_vars": spec.n,
                        "samples_per_gpu_per_step": n_chunk, "table_updates_per_sample": ftab.count_updates_per_sample(),
                        "achieved_GBs": rate * spec.n / 1e9, "frac_of_hbm_peak": rate * spec.n / 1e9 / (peak_gbs * world)}
    del ev, outs, plans, infer, tables, codes, ftab
    torch.cuda.empty_cache()
    # config 4 (fit half): 200-node card-4 partial 8-tree, counting throughput on a resident chunk
    spec = synth.random_ktree_dag()
    tables = tables_from_spec(spec, dev)
    n_chunk = args.fit_chunk
    codes = sample_network(spec, seed=1237, first=rank * n_chunk, n=n_chunk, device=dev, tables=tables)
    torch.cuda.synchronize()

    def cstep(_i):
        tables.reset_counts()
        sharding.fit_sharded(tables, codes, n_chunk)

    k = max(3, min(args.steps, 5))
    sec = timed(cstep, k, 2)
    rate = n_chunk * world * k / sec
    out["ktree200_fit"] = {"metric": "CPT-fit samples/sec", "value": rate, "unit": "samples/s", "n_vars": spec.n,
                           "samples_per_gpu_per_step": n_chunk, "family_groups": tables.count_groups(),
                           "table_updates_per_sample": tables.count_updates_per_sample(),
                           "achieved_GBs": rate * spec.n / 1e9, "frac_of_hbm_peak": rate * spec.n / 1e9 / (peak_gbs * world)}
    # config 4 (query half): 8 patterns = random target + 10 random evidence variables, 1,048,576 rows each
    import numpy as np

    infer = bind_inference(tables)
    rng = np.random.default_rng(1240)
    rows = 1 << 20
    full = sample_network(spec, seed=1241, first=rank * rows, n=rows, device=dev, tables=tables)
    t0 = time.perf_counter()
    pats = []
    for _ in range(8):
        vs = [int(v) for v in rng.choice(spec.n, size=11, replace=False)]
        plan = infer.plan(spec.names[vs[0]], [spec.names[v] for v in vs[1:]])
        pats.append((plan, full[vs[1:]].contiguous(), torch.empty((rows, plan.card_t), dtype=torch.float32, device=dev)))
    torch.cuda.synchronize()
    compile_ms = (time.perf_counter() - t0) * 1e3
    del full

    def qstep(_i):
        for plan, ev, o in pats:
            plan.run_codes(ev, rows, out=o)

    k = max(3, min(args.steps, 10))
    sec = timed(graphed(qstep), k, 3)
    alg = sum(rows * p.algorithmic_bytes_per_row() for p, _, _ in pats) * world
    out["ktree200_ve"] = {"metric": METRIC, "value": rows * len(pats) * world * k / sec, "unit": "queries/s",
                          "rows_per_gpu_per_pattern": rows, "patterns": len(pats), "achieved_GBs": alg * k / sec / 1e9,
                          "frac_of_hbm_peak": alg * k / sec / 1e9 / (peak_gbs * world), "plan_compile_ms_total": compile_ms,
                          "table_cells": [p.stats.final_tables[0][1] for p, _, _ in pats],
                          "per_row_hidden": [p.stats.per_row_hidden for p, _, _ in pats]}
    del pats, tables, codes, infer
    torch.cuda.empty_cache()
    # config 5: 1000-node layered DAG (20 x 50, cards 2..8).  (a) the 64 uniformly random patterns of the survey:
    # how many compile (their induced width is beyond exact inference); (b) 64 patterns inside the first five layers:
    # compile time and execute throughput over the ones that compile (gather plans and per-row elimination mixed)
    from continuousbayesiannetwork_b200.ve import PlanTooLarge, RowPlan

    spec = synth.layered_dag()
    tables, infer = install_cpts(spec, dev)
    rng = np.random.default_rng(1241)
    n_ok = 0
    t0 = time.perf_counter()
    for _ in range(64):
        kk = int(rng.integers(5, 51))
        vs = [int(v) for v in rng.choice(spec.n, size=kk + 1, replace=False)]
        try:
            infer.plan(spec.names[vs[0]], [spec.names[v] for v in vs[1:]])
            n_ok += 1
        except PlanTooLarge:
            pass
    random_ms = (time.perf_counter() - t0) * 1e3
    rows = args.layered_rows
    full = sample_network(spec, seed=1244, first=rank * rows, n=rows, device=dev, tables=tables)
    rng = np.random.default_rng(1242)
    pats, n_rowplans = [], 0
    t0 = time.perf_counter()
    for _ in range(64):
        kk = int(rng.integers(5, 51))
        vs = [int(v) for v in rng.choice(5 * 50, size=kk + 1, replace=False)]
        try:
            plan = infer.plan(spec.names[vs[0]], [spec.names[v] for v in vs[1:]])
        except PlanTooLarge:
            continue
        n_rowplans += isinstance(plan, RowPlan)
        pats.append((plan, full[vs[1:]].contiguous(), torch.empty((rows, plan.card_t), dtype=torch.float32, device=dev)))
    torch.cuda.synchronize()
    compile_ms = (time.perf_counter() - t0) * 1e3
    del full

    def lstep(_i):
        for plan, ev, o in pats:
            plan.run_codes(ev, rows, out=o)

    # arithmetic of the per-row schedules: multiply-adds per row, timed separately from the gather plans
    row_pats = [x for x in pats if isinstance(x[0], RowPlan)]
    row_madds = sum(p.stats.per_row_madds for p, _, _ in row_pats)

    def rstep(_i):
        for plan, ev, o in row_pats:
            plan.run_codes(ev, rows, out=o)

    row_sec = timed(graphed(rstep), 3, 1) / 3 if row_pats else 0.0
    k = 3
    sec = timed(graphed(lstep), k, 1)
    alg = sum(rows * p.algorithmic_bytes_per_row() for p, _, _ in pats) * world
    out["layered1000_ve"] = {"metric": METRIC, "value": rows * len(pats) * world * k / sec, "unit": "queries/s",
                             "uniform_random_patterns": {"attempted": 64, "compiled": n_ok, "planner_ms_total": random_ms,
                                                         "note": "induced width of the hidden part is 2^43+ cells for 63 of 64 patterns (DESIGN.md section 5)"},
                             "first_five_layers_patterns": {"attempted": 64, "compiled": len(pats), "per_row_plans": n_rowplans,
                                                            "plan_compile_ms_total": compile_ms},
                             "per_row_plans": {"plans": len(row_pats), "multiply_adds_per_row_all_plans": row_madds,
                                               "ms_per_pass": row_sec * 1e3,
                                               "achieved_Gmadd_s": (row_madds * rows / row_sec / 1e9) if row_sec else None,
                                               "note": "fp32 FMA peak of the part is ~37,000 Gmadd/s: the executor is bound by shared-memory / L1 latency per term, not by the FMA pipe"},
                             "rows_per_gpu_per_pattern": rows, "achieved_GBs": alg * k / sec / 1e9,
                             "frac_of_hbm_peak": alg * k / sec / 1e9 / (peak_gbs * world)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU)
    ap.add_argument("--fit-chunk", type=int, default=1 << 25)
    ap.add_argument("--fit-samples", type=int, default=1 << 28)
    ap.add_argument("--cpu-budget-s", type=float, default=12.0)
    ap.add_argument("--graph-streams", type=int, default=3)
    ap.add_argument("--layered-rows", type=int, default=1 << 18)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
