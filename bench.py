#!/usr/bin/env python
"""Benchmark of the per-time-step particle pipeline (BASELINE.json metric:
particle-steps/s on the 3-D SPHERIC test 2 dam break).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours      : the reference's unchanged 116-tool pipeline (resolved XML of
            examples/3D/spheric_testcase2_dambreak) driven by the C++ host
            (libaquahost.so) on the sm_100a kernels of libaquacuda.so.
reference : the CPU restatement of the same pipeline (oracle/, "port": the
            reference itself needs OpenCL/Xerces/VTK and cannot be built here)
            on all host threads, on a bounded sample of the same workload.
One JSON line is printed by rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-steps/s"
WORKLOAD = "3D SPHERIC test 2 dam break with obstacle (examples/3D/spheric_testcase2_dambreak)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class Clocks(threading.Thread):
    """nvidia-smi sampling during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_port(n_sample, steps, warmup, threads, maxiter):
    """Oracle interpreter (CPU) on the same pipeline; returns (particle-steps/s, N, ms/step)."""
    from aquagpusph_b200 import cases, casegen
    from oracle import interp, oracle as O
    O.build()
    O.set_threads(threads)
    c = cases.spheric2_dam_break(n_sample, 3.0)
    ov = {"iter_midpoint_max": maxiter} if maxiter > 0 else None
    xml = casegen.instantiate("spheric2_dambreak_3d", c, (c["N"] - 8, 8), ov)
    I = interp.Interpreter(xml, 3)
    for k in casegen.STATE_FIELDS:
        I.V[k][...] = c[k]
    for _ in range(warmup):
        I.step()
    t0 = time.time()
    for _ in range(steps):
        I.step()
    dt = time.time() - t0
    return c["N"] * steps / dt, c["N"], 1e3 * dt / steps


def run_reference(a, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    v, N, ms = cpu_port(a.cpu_n, a.steps, a.warmup, threads, a.maxiter)
    sample = ("same case generator and 116-tool pipeline at n_fluid=%d (N=%d), %d steps after %d "
              "warm-up" % (a.cpu_n, N, a.steps, a.warmup))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "particle-steps/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_particles": N, "iter_midpoint_max": a.maxiter or 30},
        "cpu_baseline": {"value": v, "unit": "particle-steps/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": "particle-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }), flush=True)


def pinned(actx, shape, dtype):
    from aquagpusph_b200 import _lib
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = ctypes.c_void_p()
    actx._chk(_lib.lib().aqc_host_alloc(actx.h, nbytes, ctypes.byref(p)))
    buf = (ctypes.c_char * nbytes).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def run_ours(a, rank, world, local_rank):
    from aquagpusph_b200 import _lib, casegen, host
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=180))
    host.set_log_level(3)
    ov = {"iter_midpoint_max": a.maxiter} if a.maxiter > 0 else None
    if world > 1 or a.slab_pipeline:
        # BASELINE config 3: the dam break cut in y slabs, one per GPU, the reference's MPI
        # example pipeline with migration + halo exchange over NCCL (weak scaling: a.n fluid
        # particles per GPU); the 128-byte NCCL id travels over torch.distributed
        uid = [host.comm_unique_id() if (rank == 0 and world > 1) else None]
        if world > 1:
            dist.broadcast_object_list(uid, src=0)
        sim, case = casegen.spheric2_slab(a.n * world, rank, world, overrides=ov,
                                          device=local_rank, unique_id=uid[0], delta_sph=not a.mpi_example)
        N = case["N"] - case["n_buffer"]      # buffer rows are not particles
        workload = ("3D SPHERIC test 2 dam break, %d y-slabs: " % world) + (
            "examples/3D/spheric_testcase2_dambreak_mpi pipeline (migration + halo over NCCL; no delta-SPH / MLS)"
            if a.mpi_example else
            "the physics of the single-GPU pipeline (examples/3D/spheric_testcase2_dambreak: delta-SPH, MLS, "
            "BIe) on the reference's MPI presets, remote delta-SPH / MLS terms added (casegen.slab_delta_sph)")
    else:
        sim, case = casegen.spheric2(a.n, overrides=ov, device=local_rank)
        N = case["N"]
        workload = WORKLOAD
    actx = _lib.Context.borrow(sim.cuda_ctx(), 3)

    def barrier():
        sim.sync()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    N_all = int(sum_over_ranks(N))

    def sweeps_done():
        # executions of the interaction sweep = inner (midpoint) iterations so far;
        # iter_midpoint itself is overwritten by the autostop preset when it converges
        return sum(n for name, n, _ in sim.tool_times() if name == "cfd interactions")

    trace = os.environ.get("AQ_BENCH_TRACE") == "1" and rank == 0
    # the timed window lies inside steps 10 ... 110 of the run, not on the column at rest
    # (the collapse has started, the midpoint solver needs its steady number of sub-iterations):
    # the state is evolved on the GPU during set-up, outside every timed region
    if a.pre_steps > 0:
        sim.step(a.pre_steps)
    for w in range(a.warmup):
        n0 = sweeps_done()
        sim.step(1)
        if trace:
            print("warm-up step %d: %d inner iterations, residual %.4g, dt %.4g"
                  % (w, sweeps_done() - n0, float(sim.scalar("Residual_midpoint")),
                     float(sim.scalar("dt"))), file=sys.stderr, flush=True)
    barrier()
    clocks = Clocks(local_rank)
    if rank == 0:
        clocks.start()
    # ---- device-resident timing: K steps between CUDA events on the stream
    e0, e1 = actx.event(), actx.event()
    l0 = sim.launch_count()

    inner = -sweeps_done()
    pc0 = actx.pairs_cache_stats()
    barrier()
    actx.record(e0)
    for _ in range(a.steps):
        sim.step(1)
    actx.record(e1)
    barrier()
    inner += sweeps_done()
    pc1 = actx.pairs_cache_stats()
    ms = max_over_ranks(actx.elapsed_ms(e0, e1))
    launches = sim.launch_count() - l0
    value = N_all * a.steps / (ms * 1e-3)
    # the clock samples belong to the device-timed region; nvidia-smi polling perturbs the
    # host-timed end-to-end loop below (driver locks), so it stops here
    clocks.stop_flag = True
    if rank == 0:
        clocks.join(timeout=10)

    # ---- end to end: host buffers in, host buffers out, every step
    fields = ["r", "u", "dudt", "rho", "drhodt", "m", "imove"]
    # the whole evolving state comes back and is fed forward: with stale rates (dudt, drhodt) in the
    # next step's input the midpoint iteration starts from an inconsistent state and needs more
    # sub-iterations than the device-resident run
    outs = ["r", "u", "rho", "p", "dudt", "drhodt"]
    if world > 1:   # particles may migrate between the slabs
        outs += ["m", "imove"]
    NA = case["N"]   # array length on this rank (includes the buffer rows of a slab)
    hin = {}
    for k in fields:
        cur = sim.download(k, np.int32 if k == "imove" else np.float32)
        hin[k] = pinned(actx, cur.shape, cur.dtype)
        hin[k][...] = cur
    hout = {k: pinned(actx, hin[k].shape if k in hin else (NA,),
                      np.int32 if k == "imove" else np.float32) for k in outs}
    h2d = sum(v.nbytes for v in hin.values())
    d2h = sum(v.nbytes for v in hout.values()) + 4
    barrier()
    inner_e2e = -sweeps_done()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        for k in fields:
            sim.upload(k, hin[k])
        sim.step(1)
        for k in outs:
            sim.download(k, hout[k].dtype, out=hout[k])
        dt_now = float(sim.scalar("dt"))
        for k in outs:
            if k in hin:   # next step starts from this step's result: swap the pinned buffers
                hin[k], hout[k] = hout[k], hin[k]
    barrier()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
    inner_e2e += sweeps_done()
    e2e = N_all * a.steps / (e2e_ms * 1e-3)
    clocks.stop_flag = True

    if dist is not None:
        dist.barrier()   # last collective: what follows is rank 0's own post-processing
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel, timed alone on the state the K steps left: the
    # fused fluid sweep (cfd/Shepard.cl + cfd/Interactions.cl + cfd/deltaSPH.cl full/lapp in
    # one pass -- the launch the pipeline's planner makes 3x per step), CUDA events on the
    # context's stream
    pk, pk_kind = peaks()
    V = {}
    delta_sph = sim.has_array("lap_p") and sim.has_array("lap_p_corr")
    for k, dt_ in (("imove", np.int32), ("r", np.float32), ("u", np.float32), ("rho", np.float32),
                   ("m", np.float32), ("p", np.float32), ("grad_p", np.float32),
                   ("lap_u", np.float32), ("div_u", np.float32), ("shepard", np.float32),
                   ("lap_p", np.float32), ("lap_p_corr", np.float32), ("icell", np.uint32),
                   ("ihoc", np.uint32)):
        if k in ("lap_p", "lap_p_corr") and not delta_sph:
            continue
        n_, eb = sim.array_info(k)
        V[k] = actx.wrap(lib_ptr(sim, k), (n_, eb // 4) if eb > 4 else (n_,), dt_)
    V["N"] = NA
    V["n_cells"] = sim.scalar("n_cells", np.uint32, 4)
    d = _lib.Defs()
    hh = float(sim.scalar("h"))
    actx.dims = 3
    n_pairs = actx.zeros(NA, np.uint32)
    V["n_pairs"] = n_pairs
    actx.launch("aqua/diag.cl", "count_pairs", V)
    pairs = int(n_pairs.get().astype(np.uint64).sum())
    members = [("cfd/Shepard.cl", "entry"), ("cfd/Interactions.cl", "entry")]
    if delta_sph:
        members += [("cfd/deltaSPH.cl", "full"), ("cfd/deltaSPH.cl", "lapp")]
    assert sim.fused_groups() >= 1, "the pipeline did not fuse the fluid sweeps"
    for _ in range(2):
        actx.launch_fused(members, V)
    k0, k1 = actx.event(), actx.event()
    reps = 5
    actx.record(k0)
    for _ in range(reps):
        actx.launch_fused(members, V)
    actx.record(k1)
    kms = actx.elapsed_ms(k0, k1) / reps
    # distinct arrays of the fused pass, 3-D, 32-bit indices: reads imove 4, r 16,
    # u 16, rho 4, m 4, p 4, icell 4; writes grad_p 16, lap_u 16, div_u 4, shepard 4, lap_p 4,
    # lap_p_corr 16 = 112 B/particle (the four stand-alone sweeps: 88 + 36 + 52 + 40 = 216)
    alg_bytes = (112.0 if delta_sph else 92.0) * NA
    # SURVEY 8(d) per true pair: Interactions 52 flop; with the shared |r_ij|, q and F(q) the
    # Shepard, full and lapp members add 6 + 6 + 2
    alg_flops = (66.0 if delta_sph else 58.0) * pairs
    achieved = alg_bytes / (kms * 1e-3) / 1e9
    pcs = actx.pairs_cache_stats()
    cached = pcs["bytes"] > 0 and pc1["hits"] > pc0["hits"]
    lists = cached and os.environ.get("AQC_PAIR_LISTS", "1") != "0"
    # DRAM traffic of one launch: NOT measured in this run -- a constant cited from the ncu capture
    # of the same kernel on the same workload (ncu --set full; dram__bytes_read + dram__bytes_write)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "fused_traffic.json")
    if os.path.exists(tp):
        rec = json.load(open(tp))
        if rec.get("n_particles") == NA and rec.get("lists", False) == lists:
            traffic = rec.get("dram_bytes_per_launch")
            traffic_src = "cited from " + rec.get("source", tp) + " (not measured in this run)"
    # measured FP32 (non-tensor) peak of THIS device: FMA micro-benchmark (csrc/peak.cu), timed here
    ffma, ffma2 = actx.fp32_peak()
    tflops = alg_flops / (kms * 1e-3) / 1e12
    kname = ("sweep4_kernel<PFusedFluid<3,...>> reading the neighbour lists (" if lists else
             "sweep3_kernel<PFusedFluid<3,...>, 2, 8> reading the pair masks (" if cached else
             "sweep3_kernel<PFusedFluid<3,...>, 0, 8> (")
    roof = {"bound": "fp32",
            "kernel": kname + " + ".join(m[0] + "::" + m[1] for m in members) + ")",
            "achieved": tflops, "peak": ffma, "unit": "TFLOP/s", "frac": tflops / ffma,
            "peak_kind": "measured in this run (aqc_fp32_peak: scalar FFMA chains; packed FFMA2 %.1f)" % ffma2,
            "traffic": traffic, "traffic_source": traffic_src,
            "ms_per_launch": kms, "algorithmic_flops": alg_flops, "pairs": pairs,
            "launches_per_step": inner / a.steps,
            "note": "neighbour sweeps are bound by FP32 issue and the shared-memory pipe, not by HBM "
                    "(>500 flop/B against a ridge of ~11, SURVEY 8(d)); no tensor-core path exists "
                    "for this non-contraction work",
            # the same launch against the HBM roof, and what it streams beyond the particle arrays
            "hbm": {"algorithmic_bytes": alg_bytes, "achieved": achieved, "peak": pk["hbm_gbs"],
                    "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "peak_kind": pk_kind,
                    "pair_cache_stream_bytes": pcs["bytes"] if cached else 0}}
    # ---- the HBM-bound stages north_star names, timed alone on the same state (CUDA events):
    # the link-list build (min/max, icell, stable sort of (icell, id), heads: 84 B + 4 n_cells.w / N per
    # particle) and the permutation of the 11 particle fields (212 B per particle)
    if world > 1:
        # (the link-list build of a slab all-reduces r_min / r_max: rank 0 cannot run it alone once
        # the peers have left; these two stages are per-GPU work and are reported by the N = 1 line)
        roof["stages"] = {"note": "per-GPU stages, timed by the N = 1 line"}
    else:
        try:
            roof["stages"] = hbm_stages(sim, actx, V, NA, pk["hbm_gbs"])
        except Exception as e:   # never lose the line over the extra measurement
            roof["stages"] = {"error": str(e)[:200]}
    del d, hh
    # ---- N > 1 runs the physics of the N = 1 headline (delta-SPH, MLS, BIe) on slabs: the
    # reference's MPI presets plus the remote delta-SPH / MLS terms its own MPI preset lacks
    # (casegen.slab_delta_sph; --mpi-example gives the reference's lighter 131-tool example instead).
    # On one rank that pipeline is the 116-tool one bit for bit, so value(N) / (N value(1)) of the
    # driver is a weak-scaling efficiency; rank 0 still times ONE slab of the same per-GPU size
    # through the very same tool list on its GPU, after the other ranks have left.
    same1 = None
    if world > 1 and not a.skip_same_pipeline:
        try:
            sim1, case1 = casegen.spheric2_slab(a.n, 0, 1, overrides=ov, device=local_rank,
                                                unique_id=None, delta_sph=not a.mpi_example)
            ctx1 = _lib.Context.borrow(sim1.cuda_ctx(), 3)
            sim1.step(a.pre_steps)
            for _ in range(a.warmup):
                sim1.step(1)
            sim1.sync()
            s0, s1 = ctx1.event(), ctx1.event()
            ctx1.record(s0)
            for _ in range(a.steps):
                sim1.step(1)
            ctx1.record(s1)
            sim1.sync()
            ms1 = ctx1.elapsed_ms(s0, s1)
            n1 = case1["N"] - case1["n_buffer"]
            same1 = {"n_gpus": 1, "n_particles": n1, "ms_per_step": ms1 / a.steps,
                     "value": n1 * a.steps / (ms1 * 1e-3),
                     "note": "same MPI-example pipeline and per-GPU size on one GPU: the "
                             "denominator of the weak-scaling efficiency of this line"}
            same1["weak_scaling_efficiency"] = value / (world * same1["value"])
        except Exception as e:   # never lose the line over the extra measurement
            same1 = {"error": str(e)[:200]}
    # ---- CPU baseline (oracle port), bounded sample: on rank 0 of the N = 1 run only
    threads = os.cpu_count() or 1
    cpu_baseline = None
    if world == 1:
        cv, cN, cms = cpu_port(a.cpu_n, a.cpu_steps, 1, threads, a.maxiter)
        cpu_baseline = {"value": cv, "unit": "particle-steps/s", "cores": threads, "kind": "port",
                        "sample": "same pipeline at n_fluid=%d (N=%d), %d steps after 1 warm-up, "
                                  "%.0f ms/step" % (a.cpu_n, cN, a.cpu_steps, cms)}
    line = {
        "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "n_particles": N_all, "n_particles_rank0": N,
                   "n_fluid": case.get("n_fluid_global", case["n_fluid"]),
                   "hfac": 3.0, "pipeline_tools": len(sim.tools()),
                   "mean_inner_iterations": inner / a.steps,
                   "first_timed_step": a.pre_steps + a.warmup,
                   "iter_midpoint_max": a.maxiter or 30,
                   # `while` loops run as a CUDA graph WHILE node from their second pass on
                   # (host/devloop.hpp; AQUA_DEVICE_LOOPS=0 keeps them on the host), and the passes
                   # made there since the set-up
                   "device_loops": {"loops": sim.device_loops(),
                                    "runs_and_passes": list(sim.device_loop_stats()),
                                    "tools_on_second_stream": sim.device_loop_branch_tools(),
                                    "last_step_host_cost": sim.device_loop_timing()},
                   "l2": "inputs larger than L2 (%.0f MB of particle arrays)" % (560.0 * N / 1e6),
                   "multi_gpu": ("y-slab decomposition, mpi-sync over NCCL send/recv (halo of r, u, rho, m and of "
                                 "the delta-SPH gradient every sub-iteration), dt and residual all-reduced")
                   if world > 1 else "single GPU",
                   "one_gpu_same_pipeline": same1,
                   # neighbour sweeps that read the hit masks of one builder pass instead of
                   # filtering their candidates again (include/aquacuda.h, aqc_pairs_cache_*)
                   "pair_mask_cache": {"builds_per_step": (pc1["builds"] - pc0["builds"]) / a.steps,
                                       "sweeps_served_per_step": (pc1["hits"] - pc0["hits"]) / a.steps,
                                       "device_bytes": pc1["bytes"]}},
        "e2e": {"value": e2e, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / a.steps,
                "mean_inner_iterations": inner_e2e / a.steps},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": roof,
        "cpu_baseline": cpu_baseline,   # (timed by the N = 1 line)
    }
    print(json.dumps(line), flush=True)
    _ = dt_now
    if dist is not None:
        dist.destroy_process_group()


def hbm_stages(sim, actx, V, NA, peak):
    """Link-list build and field permutation on the state the K steps left, each timed alone."""
    out = {}
    dims = 3

    def wrap(name, dt_):
        n_, eb = sim.array_info(name)
        return actx.wrap(lib_ptr(sim, name), (n_, eb // 4) if eb > 4 else (n_,), dt_)

    r_in = wrap("r_in", np.float32)
    icell, perm, inv = (actx.empty(NA, np.uint32) for _ in range(3))
    h = float(sim.scalar("h"))
    ihoc = None
    for rep in range(8):
        if rep == 3:
            e0, e1 = actx.event(), actx.event()
            actx.record(e0)
        _, _, nc, ihoc = actx.linklist(r_in, 2.0, h, icell, ihoc, perm, inv)
    actx.record(e1)
    ms = actx.elapsed_ms(e0, e1) / 5
    b = (84.0 + 4.0 * float(nc[3]) / NA) * NA
    out["linklist"] = {"ms": ms, "algorithmic_bytes": b, "achieved": b / (ms * 1e-3) / 1e9,
                       "peak": peak, "unit": "GB/s", "frac": b / (ms * 1e-3) / 1e9 / peak}
    # basic/Sort.cl stage1 + stage2 (+ the six backup copies of basic.xml:141-148 are not algorithmic)
    W = dict(V)
    for k, dt_ in (("id", np.uint32), ("iset", np.uint32), ("normal", np.float32), ("tangent", np.float32),
                   ("dudt", np.float32), ("drhodt", np.float32), ("id_sorted", np.uint32)):
        W[k] = wrap(k, dt_)
    for k, dt_ in (("id", np.uint32), ("iset", np.uint32), ("imove", np.int32), ("r", np.float32),
                   ("normal", np.float32), ("tangent", np.float32), ("rho", np.float32), ("m", np.float32),
                   ("u", np.float32), ("dudt", np.float32), ("drhodt", np.float32)):
        W[k + "_in"] = wrap(k + "_in", dt_)
    for rep in range(8):
        if rep == 3:
            actx.record(e0)
        actx.launch("basic/Sort.cl", "stage1", W)
        actx.launch("basic/Sort.cl", "stage2", W)
    actx.record(e1)
    ms = actx.elapsed_ms(e0, e1) / 5
    b = 212.0 * NA
    out["permutation"] = {"ms": ms, "algorithmic_bytes": b, "achieved": b / (ms * 1e-3) / 1e9,
                          "peak": peak, "unit": "GB/s", "frac": b / (ms * 1e-3) / 1e9 / peak}
    return out


def lib_ptr(sim, name):
    from aquagpusph_b200 import host
    return host.lib().aqh_array_devptr(sim.h, name.encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", "--n", dest="n", type=int, default=1000000,
                    help="fluid particles per GPU (Create.py n)")
    ap.add_argument("--cpu-n", type=int, default=130000, help="fluid particles of the CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=3, help="timed steps of the cpu_baseline leg")
    ap.add_argument("--pre-steps", type=int, default=30,
                    help="steps evolved on the GPU before the warm-up (outside the timed regions)")
    ap.add_argument("--slab-pipeline", action="store_true",
                    help="run the multi-GPU (slab) pipeline also on one GPU, for scaling studies")
    ap.add_argument("--mpi-example", action="store_true",
                    help="N > 1: the reference's MPI example pipeline (no delta-SPH / MLS) instead of the "
                         "single-GPU pipeline's physics on slabs")
    ap.add_argument("--skip-same-pipeline", action="store_true",
                    help="N > 1: do not time the one-GPU run of the same slab pipeline afterwards")
    ap.add_argument("--maxiter", type=int, default=0, help="pin iter_midpoint_max (0: case default 30)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.warmup < 3 and a.impl == "ours":
        a.warmup = 3
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    try:
        run_ours(a, rank, world, local_rank)
    except BaseException:
        # a rank that failed must END the job: the peers' waits are bounded (aqc_comm_wait), and the
        # launcher kills them as soon as this process is gone -- no destructors, no barriers
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        sys.stdout.flush()
        os._exit(1)


if __name__ == "__main__":
    main()
