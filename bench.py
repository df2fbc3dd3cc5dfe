#!/usr/bin/env python
"""bench.py - witnesses/sec + R1CS constraints checked/sec for registerIdentity (SHA-256 + RSA-2048).

One "step" = one pass of the hot path (evaluate every signal of every lane, check every constraint, fold every
wire of every lane into the witness digest) over one batch of synthetic passports of the north-star circuit
RegisterIdentityBuilder(1,256,3,4,600,248,1,1496,3,256) (/root/reference/hardhat.config.ts:29).

  python bench.py [--gpus N --steps K --warmup W] [--batch B]      our arm (CUDA, one rank per GPU)
  python bench.py --impl reference ...                             CPU arm: the oracle evaluator on host cores
  python bench.py --sweep [--gpus N]                               BASELINE config 5: 1 Ki .. 1 Mi GLOBAL passports

Both arms run the SAME compilation (`c3`: alias / truth-table / symbolic / bit-view rows are discharged by
compile-time proofs, the rest is evaluated at run time) and produce the full witness: our arm folds all
2 251 704 wires of every lane into the per-lane digest on the device (pzk.h "witness digest"), the CPU arm
materialises the vector.

`value`     = whole-job witnesses/s, inputs resident in HBM (CUDA events on the library's stream, max over ranks);
`e2e`       = the same through the C ABI call `pzk_witness_batch_packed_digest` with pinned HOST buffers: H2D of the
              packed records, D2H of status + first failing constraint + public signals + digest inside the timed
              region (tiles of one wave, copies of the neighbouring tiles under the kernels);
`roofline`  = the evaluator against the INTEGER pipe: algorithmic 32-bit multiply-adds (136 per Fr
              product + 1 per narrow record) per launch / CUDA-event launch time, against the IMAD rate measured in
              this run by tests/cuda/imad_peak.cu; `roofline.hbm` = measured DRAM bytes (ncu capture of this
              build, profiles/) over the same time against the measured copy bandwidth;
`verdict_only` = the same batch without the digest (values that never leave the operand cache are not stored);
`all_rows_pair` = GPU and CPU both on `c3_allrows` (every non-alias row evaluated at run time, no views);
`r1cs_pipeline` = the north star's two kernels back to back: evaluate -> device-resident hand-off -> every row of
              the .r1cs streamed through TMA on the selected lanes.
The batch shards across ranks with no collective: scaling is weak; --sweep is the
strong-scaling view (fixed global batch split with sharding.shard_bounds).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "witnesses/sec + R1CS constraints checked/sec, registerIdentity SHA256/RSA2048"
WORKLOAD = "registerIdentity_1_256_3_4_600_248_1_1496_3_256 (SHA-256 + RSA-2048 e=65537), synthetic passports"
UNIQUE = 256  # distinct signed passports generated on the host; tiled to fill the batch
PROGRAM = "c3"


def make_inputs(meta, batch, seed):
    import numpy as np
    from passport_zk_circuits_b200 import witness as W
    from passport_zk_circuits_b200.passports import C3, PassportFactory
    fac = PassportFactory(C3, seed=seed, n_sig_keys=4, n_aa_keys=4)
    n = min(UNIQUE, batch)
    uniq = W.pack_inputs_fast(meta, [fac.make(i).inputs for i in range(n)])
    reps = (batch + n - 1) // n
    return np.tile(uniq, (reps, 1, 1))[:batch].copy()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def _ref_worker(args):
    prog, inputs, check = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref as oracle_ref
    rp = oracle_ref.RefProgram(prog)
    t = time.time()
    bad = 0
    for row in inputs:
        st, fb, _ = rp.witness(row, want_witness=True, check_rows=check)   # the whole vector, like calculateWitness
        bad += st != 0
    return time.time() - t, bad


def cpu_reference(prog, inputs, workers, per_worker):
    """The oracle evaluator (oracle/ssa_ref.c) on `workers` host processes, `per_worker` witnesses each.
    Returns witnesses/s (wall clock of the slowest worker)."""
    import multiprocessing as mp
    jobs = [(prog, inputs[(w * per_worker) % len(inputs):][:per_worker], True) for w in range(workers)]
    jobs = [(p, i if len(i) == per_worker else inputs[:per_worker], c) for p, i, c in jobs]
    t0 = time.time()
    with mp.get_context("spawn").Pool(workers) as pool:
        res = pool.map(_ref_worker, jobs)
    wall = time.time() - t0
    slow = max(r[0] for r in res)
    assert sum(r[1] for r in res) == 0, "oracle evaluator reported failing lanes on valid passports"
    return workers * per_worker / slow, wall


def measured_integer_peak():
    """tests/cuda/imad_peak.cu on this GPU, now: dependent-free IMAD rate and register-operand Montgomery products."""
    exe = os.path.join(ROOT, "tests", "bin", "imad_peak")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:  # the roofline then says so instead of inventing a peak
        return {"error": f"{type(e).__name__}: {e}"}


def ncu_dram_bytes():
    """DRAM bytes per launch from the committed ncu captures of this build (profiles/r2_eval_dram.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_eval_dram.json")))
    except Exception:
        return None


def rows_of(stats, n_constraints):
    return {"total": n_constraints, "static_alias": stats["static_rows"], "static_table_proof": stats.get("table_rows", 0),
            "static_symbolic_proof": stats.get("symbolic_rows", 0), "static_bit_view_proof": stats.get("view_rows", 0),
            "static_definitional": stats.get("def_rows", 0),
            "runtime": stats["i64_rows"] + stats["int_rows"] + stats["field_rows"] + stats.get("range_rows", 0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0, help="passports per GPU per step (0 = two waves of resident CTAs)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="witnesses per worker for the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip verdict_only / all_rows_pair / r1cs_pipeline")
    ap.add_argument("--r1cs-lanes", type=int, default=512, help="lanes handed to the R1CS stream kernel (0 = skip)")
    ap.add_argument("--sweep", action="store_true", help="BASELINE config 5: global batches 1 Ki .. 1 Mi split over the ranks")
    ap.add_argument("--sweep-sizes", default="1024,4096,16384,65536,262144,1048576")
    ap.add_argument("--sweep-out", default="", help="append one JSON object per sweep point to this file")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from passport_zk_circuits_b200 import witness as W
    import numpy as np

    prog = W.artifact(PROGRAM)

    # ------------------------------------------------------------------ reference arm (CPU)
    if a.impl == "reference":
        if rank != 0:
            return
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref as oracle_ref
        oracle_ref.build()
        rp = oracle_ref.RefProgram(prog)
        cores = os.cpu_count() or 1
        per = a.cpu_sample or 128
        inputs = make_inputs(rp.meta, min(UNIQUE, cores * per), seed=1)
        for _ in range(max(a.warmup, 0) and 1):
            cpu_reference(prog, inputs, cores, 2)
        vals = []
        for _ in range(a.steps):
            v, _ = cpu_reference(prog, inputs, cores, per)
            vals.append(v)
        v = statistics.mean(vals)
        stats = rp.meta["stats"]
        rows = rows_of(stats, rp.n_constraints)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "witnesses/s", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000.0 * cores * per / v,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64+fr256",
                "data": "synthetic", "constraints_discharged_per_sec": v * rp.n_constraints,
                "constraints_evaluated_per_sec": v * rows["runtime"],
                "config": {"workload": WORKLOAD, "program": PROGRAM, "sample": f"{cores * per} witnesses per step"},
                "cpu_baseline": {"value": v, "unit": "witnesses/s", "cores": cores, "kind": "port",
                                 "sample": f"{cores} processes x {per} witnesses per step (oracle/ssa_ref.c on the "
                                           f"program `{PROGRAM}` our arm runs, full witness materialised)"},
                "e2e": {"value": v, "unit": "witnesses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm (CUDA)
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    calc = W.WitnessCalculator(prog, device=local_rank)
    calc.set_digest(True)
    n_pub = calc.n_public
    L = calc._L

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def pinned(B, packed_src):
        buf = {"in": torch.empty(packed_src.shape, dtype=torch.uint8, pin_memory=True),
               "status": torch.empty(B, dtype=torch.int32, pin_memory=True),
               "bad": torch.empty(B, dtype=torch.int64, pin_memory=True),
               "pub": torch.empty((B, n_pub, 4), dtype=torch.int64, pin_memory=True),
               "digest": torch.empty((B, 4), dtype=torch.int64, pin_memory=True)}
        buf["in"].numpy()[...] = packed_src
        return buf

    def e2e_call(c, buf, B, digest=True):
        rc = L.pzk_witness_batch_packed_digest(c._h, buf["in"].data_ptr(), B, buf["status"].data_ptr(), buf["bad"].data_ptr(),
                                               buf["pub"].data_ptr(), buf["digest"].data_ptr() if digest else None)
        assert rc == 0, (rc, L.pzk_last_error(c._h))

    unique_inputs = make_inputs(calc.meta, UNIQUE, seed=1 + (0 if a.sweep else rank))
    packed_unique = calc.pack(unique_inputs)

    # ------------------------------------------------------------------ config 5: batch-size sweep, strong scaling
    if a.sweep:
        from passport_zk_circuits_b200.sharding import gather_lanes, shard_bounds
        points = []
        for G in [int(x) for x in a.sweep_sizes.split(",")]:
            lo, hi = shard_bounds(G, rank, world)
            B = hi - lo
            idx = (np.arange(lo, hi) % len(packed_unique))
            buf = pinned(max(B, 1), packed_unique[idx] if B else packed_unique[:1])
            reps = 3 if G <= 65536 else 2
            if B:
                e2e_call(calc, buf, B)          # warm-up (allocations, first touch)
            barrier()
            t0 = time.time()
            for _ in range(reps):
                if B:
                    e2e_call(calc, buf, B)
            barrier()
            dt = (time.time() - t0) / reps
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            status = gather_lanes(buf["status"].numpy()[:B].copy(), G, dist)
            digest0 = gather_lanes(buf["digest"].numpy()[:B].copy(), G, dist)
            assert int((status != 0).sum()) == 0 and digest0.shape == (G, 4)
            pt = {"config": 5, "global_batch": G, "n_gpus": world, "per_gpu_batch": [int(shard_bounds(G, r, world)[1] - shard_bounds(G, r, world)[0]) for r in range(world)],
                  "ms": 1000.0 * float(tt.item()), "witnesses_per_s": G / float(tt.item()), "unit": "witnesses/s",
                  "wave_lanes": int(calc.wave_lanes()), "waves_per_gpu": B / calc.wave_lanes(),
                  "mode": "end to end: pinned host records in, status + first_bad + public + digest out, every wire digested"}
            points.append(pt)
            if rank == 0 and a.sweep_out:
                with open(a.sweep_out, "a") as f:
                    f.write(json.dumps(pt) + "\n")
        if rank == 0:
            print(json.dumps({"metric": METRIC, "sweep": points, "n_gpus": world, "scaling": "strong",
                              "config": {"workload": WORKLOAD, "program": PROGRAM}}))
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ headline
    stats = calc.meta["stats"]
    B = a.batch or 2 * calc.wave_lanes()   # tiles are whole waves of resident CTAs
    reps = (B + len(packed_unique) - 1) // len(packed_unique)
    packed = np.tile(packed_unique, (reps, 1))[:B].copy()
    h2d = packed.nbytes
    d2h = B * (4 + 8 + n_pub * 32 + 32)

    # device-resident measurement
    calc.upload_packed(packed)
    for _ in range(a.warmup):
        calc.run(True)
    res = calc.download()
    dig = calc.download_digest()
    assert (res.status == 0).all(), "valid synthetic passports must satisfy every constraint"
    calc.profile(enable=True, reset=True)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.time()
    for _ in range(a.steps):
        calc.run(True)
    barrier()
    wall = time.time() - t0
    clocks = sampler.stop()
    prof = calc.profile()
    dev_ms = prof["run"][0]
    seg_launches = prof["eval"][1]
    calc.profile(enable=False)

    # end to end through the C ABI with pinned host buffers
    buf = pinned(B, packed)
    e2e_call(calc, buf, B)
    barrier()
    t1 = time.time()
    e2e_steps = max(1, min(a.steps, 3))
    for _ in range(e2e_steps):
        e2e_call(calc, buf, B)
    barrier()
    e2e_wall = time.time() - t1
    assert int((buf["status"].numpy() != 0).sum()) == 0
    assert np.array_equal(buf["digest"].numpy().view(np.uint64), dig), "streamed and resident runs disagree on the digest"

    times = torch.tensor([dev_ms / 1000.0, e2e_wall, wall], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_s, e2e_s, wall_s = [float(x) for x in times.tolist()]
    total = B * world
    value = total * a.steps / dev_s
    e2e_value = total * e2e_steps / e2e_s

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, peak_src = (peaks.get("hbm_gbs"), "MEASURED_PEAKS.json") if peaks.get("hbm_gbs") else (6650.0, "fallback B200_PROFILING.md")
        hist = W.program_histogram(prog)
        ipk = measured_integer_peak()
        rows = rows_of(stats, calc.n_constraints)
        eval_ms, eval_launches = prof["eval"]
        lanes_per_launch = min(B, calc.tile_lanes())
        # algorithmic IMAD of the whole timed region / CUDA-event time of its evaluator launches (eval_kernel + bjj_kernel;
        # with the fused digest the folds are inside those launches and inside that time, but not in the numerator)
        avg_launch_s = eval_ms / 1e3 / max(1, eval_launches)
        achieved = hist["algorithmic_imad"] * B * a.steps / (eval_ms / 1e3)
        imad_per_launch = achieved * avg_launch_s
        imad_peak = ipk.get("imad_per_s")
        dram = ncu_dram_bytes()
        traffic = None
        hbm = None
        if dram and dram.get("bytes_per_wave_launch_mean"):
            traffic = dram["bytes_per_wave_launch_mean"] * lanes_per_launch / dram["lanes"]
            hbm = {"achieved": traffic / avg_launch_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                   "frac": traffic / avg_launch_s / 1e9 / hbm_peak, "peak_source": peak_src, "source": dram.get("source")}
        roofline = {"bound": "imad", "kernel": "eval_kernel", "achieved": achieved / 1e9, "peak": imad_peak / 1e9 if imad_peak else None,
                    "unit": "GIMAD/s", "frac": achieved / imad_peak if imad_peak else None, "traffic": traffic,
                    "peak_source": "tests/cuda/imad_peak.cu run inside this bench (dependent-free mad.lo.u32, 148 SMs)",
                    "algorithmic_imad_per_witness": hist["algorithmic_imad"],
                    "fr_products_per_witness": hist["fr_products"], "narrow_records_per_witness": hist["narrow_records"],
                    "definition": "136 IMAD per Fr product (explicit F_MUL records + products inside the hint intrinsics + 1 per "
                                  "quadratic field row) + 4 la lb per Z-class integer product of la x lb 64-bit limbs + 1 per narrow "
                                  "record; Montgomery conversions, address arithmetic and the digest are NOT counted",
                    "z_mul_records_per_witness": hist["z_mul_records"], "z_mul_imad_per_witness": hist["z_mul_imad"],
                    "imad_per_launch": imad_per_launch, "avg_launch_ms": 1e3 * avg_launch_s, "launches": eval_launches,
                    "share_of_step": eval_ms / max(1e-9, prof["run"][0]),
                    "fr_mul_microbenchmark": {"fr_mul_per_s": ipk.get("fr_mul_per_s"),
                                              "frac_of_it": hist["fr_products"] * value / world / ipk["fr_mul_per_s"] if ipk.get("fr_mul_per_s") else None,
                                              "note": "Montgomery products per second with register operands (same fr_mul)"},
                    "integer_peak": ipk, "hbm": hbm}
        line = {"metric": METRIC, "value": value, "unit": "witnesses/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": 1000.0 * dev_s / a.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u64+fr256", "data": "synthetic",
                "constraints_discharged_per_sec": value * calc.n_constraints,
                "constraints_evaluated_per_sec": value * rows["runtime"],
                "wires_digested_per_sec": value * calc.n_wires,
                "config": {"workload": WORKLOAD, "program": PROGRAM, "batch_per_gpu": B, "global_batch": total,
                           "unique_passports_per_gpu": min(UNIQUE, B), "tile_lanes": calc.tile_lanes(),
                           "n_wires": calc.n_wires, "n_constraints": calc.n_constraints, "witness_digest": "on (every wire of every lane)",
                           "cache": "slot planes are GBs per tile (>> 126 MB L2); no flush needed",
                           "parallelism": f"batch sharded over {world} GPU(s), no collective"},
                "e2e": {"value": e2e_value, "unit": "witnesses/s", "h2d_bytes_per_step": h2d * world,
                        "d2h_bytes_per_step": d2h * world, "steps": e2e_steps,
                        "inputs": "packed records (bits as bytes, limbs as u64, field elements as 32 B) in pinned host memory",
                        "returns": "status + first failing constraint + 5 public signals + 256-bit witness digest per passport"},
                "gpu_launches": int(sum(prof[k][1] for k in ("eval", "check", "export", "digest"))),
                "kernel_ms": {k: prof[k][0] for k in ("eval", "digest", "export", "run")},
                "wall_s": wall_s, "clocks": clocks, "roofline": roofline, "rows": rows}
        if not a.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import ref as oracle_ref
            oracle_ref.build()
            cores = os.cpu_count() or 1
            per = a.cpu_sample or 128
            v, cpu_wall = cpu_reference(prog, unique_inputs, cores, per)
            line["cpu_baseline"] = {"value": v, "unit": "witnesses/s", "cores": cores, "kind": "port",
                                    "sample": f"{cores} processes x {per} witnesses (oracle/ssa_ref.c on the same program `{PROGRAM}`, "
                                              f"same inputs, full witness materialised; the reference's wasm calculator cannot run "
                                              f"on this host - no node / circom)", "seconds": cpu_wall}
        if world == 1 and not a.no_extras:
            calc.set_digest(False)
            calc.run(True)
            lean = calc.download()
            assert np.array_equal(lean.status, res.status) and np.array_equal(lean.public, res.public)
            calc.profile(enable=True, reset=True)
            torch.cuda.synchronize()
            for _ in range(2):
                calc.run(True)
            torch.cuda.synchronize()
            lms = calc.profile()["run"][0] / 2
            calc.profile(enable=False)
            t2 = time.time()
            e2e_call(calc, buf, B, digest=False)
            torch.cuda.synchronize()
            l_e2e = time.time() - t2
            line["verdict_only"] = {"value": B / (lms / 1e3), "e2e": B / l_e2e, "unit": "witnesses/s", "ms_per_step": lms,
                                    "note": "digest off: status, first_bad and public signals only; values whose every reader hits "
                                            "the operand cache are never stored (the round-1 headline mode)"}
            if a.r1cs_lanes > 0:
                # the north star's second kernel fed by the first one on the device
                try:
                    calc.set_tile_lanes(calc.wave_lanes())            # one wave of slot planes: room for the full witnesses
                    r = W.R1cs(W.artifact_r1cs(PROGRAM), local_rank)
                    n = min(a.r1cs_lanes, B)
                    lanes = np.arange(n, dtype=np.uint64) * (B // n)
                    r.check_circuit(calc, lanes[:32])                 # warm-up
                    ok, fb, t_eval, t_check = r.check_circuit(calc, lanes)
                    assert ok.all()
                    gbs = r.n_terms * 32 * n / (t_check / 1e3) / 1e9
                    rd = (ncu_dram_bytes() or {}).get("r1cs_stream")
                    line["r1cs_pipeline"] = {
                        "kernel": "r1cs_stream_kernel", "lanes_checked": n, "of_batch": B, "eval_and_export_ms": t_eval, "check_ms": t_check,
                        "witness_checks_per_s": n / (t_check / 1e3), "constraints_evaluated_per_s": n * r.n_constraints / (t_check / 1e3),
                        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                     "algorithmic_bytes_per_witness": r.n_terms * 32, "traffic": rd,
                                     "note": "algorithmic bytes = 32 B x terms; `traffic` = measured DRAM bytes per launch (ncu) when a capture of this build is committed"},
                        "note": "evaluator -> canonical blocked planes on the device (no host trip of the 72 MB witnesses) -> all "
                                f"{r.n_constraints} rows streamed with cp.async.bulk + mbarrier; a full witness is 72 MB, so the hand-off "
                                "audits a sample of the batch; the fused rows cover every lane"}
                    r.close()
                except W.PzkError as e:
                    line["r1cs_pipeline"] = {"unavailable": str(e)}
            calc.close()
            try:
                vprog = W.artifact("c3_allrows")
                v = W.WitnessCalculator(vprog, device=local_rank)
                v.upload_packed(packed)
                v.run(True)
                vres = v.download()
                assert np.array_equal(vres.status, res.status) and np.array_equal(vres.public, res.public)
                assert np.array_equal(vres.first_bad, res.first_bad)
                v.profile(enable=True, reset=True)
                torch.cuda.synchronize()
                v.run(True)
                torch.cuda.synchronize()
                vms = v.profile()["run"][0]
                vs = v.meta["stats"]
                pair = {"gpu": B / (vms / 1e3), "unit": "witnesses/s", "program": "c3_allrows", "rows": rows_of(vs, v.n_constraints),
                        "op_records_per_witness": vs["op_records"],
                        "note": "alias proofs only: every other row is evaluated at run time on both sides (no views, no digest)"}
                v.close()
                if not a.no_cpu_baseline:
                    cv, _ = cpu_reference(vprog, unique_inputs, os.cpu_count() or 1, 2)
                    pair["cpu"] = cv
                    pair["ratio"] = pair["gpu"] / cv
                line["all_rows_pair"] = pair
            except W.PzkError as e:
                line["all_rows_pair"] = {"unavailable": str(e)}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
