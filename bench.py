#!/usr/bin/env python
"""
Benchmark of the threshold-Paillier hot path on B200 (BASELINE.json metric: threshold
decrypts/sec at 2048-bit N).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

Workload (BASELINE.json configs[1]): 3 parties, corruption threshold t=1, key_length 2048
(exact 2048-bit N, 4096-bit N^2; synthetic dealer-generated key and uniformly random ciphertext
units): one *step* = one batch of B ciphertexts per GPU taken through the whole decryption:
d+1 = 3 partial decryptions  c^(e_i) mod N^2  (signed ~4190-bit per-key exponents, one of them
negative => batched modular inversion) + one share combination  (prod mod N^2, L-function,
* theta^-1 mod N).  All d+1 partials are materialised, as they are API-visible values in the
reference (paillier_shared_key.py:52-127, distributed_keygen.py:430-517).

Printed JSON line (rank 0): `value` = threshold decrypts/s with inputs resident in HBM, `e2e` =
the same through the public host-buffer API (pinned host -> device copies and result read-back
inside the timed region), `roofline` = the modexp kernel against the measured integer-multiplier
peak, `cpu_baseline` = GMP mpz_powm (the function gmpy2.powmod wraps) on all host cores.
Multi-GPU: one process per GPU (torchrun), ciphertext batches sharded by index, no data-path
collective; only a barrier and a max-over-ranks of the elapsed time go through NCCL.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "threshold decrypts/sec (2048-bit N)"
UNIT = "decrypts/s"
KEY_NAME = "cfg2_k2048_p3_t1_exact"
WORKLOAD = "cfg2: 3 parties t=1 key_length=2048 (exact 2048-bit N): 3 partial decryptions + share combination per ciphertext"


class KeyData:
    """The synthetic dealer key of the workload (tests/golden/dealer_vectors.json), as plain integers:
    n, parties, t, theta and one Shamir share of lambda*beta per party."""

    def __init__(self, d: dict) -> None:
        self.json = d
        self.n = int(d["n"], 16)
        self.parties = int(d["parties"])
        self.t = int(d["t"])
        self.theta = int(d["theta"], 16)
        self.shares = {int(i): int(v, 16) for i, v in d["shares"].items()}


def load_key() -> KeyData:
    with open(os.path.join(ROOT, "tests", "golden", "dealer_vectors.json")) as fh:
        data = json.load(fh)
    return KeyData(data["keys"][KEY_NAME]["key"])


def oracle_key(dk: KeyData):
    """The oracle's view of the same key: only the CPU-baseline / reference arm may use it."""
    from oracle import keys as okeys

    return okeys.dealer_key_from_json(dk.json)


def hbm_side(traffic_bytes: float, launch_ms: float) -> dict:
    """DRAM bytes of one launch / its duration against the measured copy bandwidth of this pool's
    B200s (MEASURED_PEAKS.json, driver-written; fallback: the profiling recipe's 6.5 TB/s)."""
    peak, src = 6500.0, "fallback 6.5 TB/s (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peak, src = float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    achieved = traffic_bytes / (launch_ms * 1e-3) / 1e9
    return {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": src}


def canonical_modexp_macs(exp_bits: int, limbs: int) -> float:
    """modmul(L) = 2L^2 + L wide-MACs; modexp(E, L) = (E + ceil(E/5) + 32)
    modmuls (squarings counted as multiplies, canonical window 5)."""
    return float(exp_bits + (exp_bits + 4) // 5 + 32) * float(2 * limbs * limbs + limbs)


def actual_modexp_macs(info: dict) -> float:
    """Wide multiply-accumulates the kernels really execute per exponentiation (block products of
    K x K limbs, low-half quotient products), from the context's shape and window parameters."""
    # `windows` multiplications in the main loop (the first one is a table load), two domain
    # conversions, one squaring per exponent bit below the first window (bounded by the bit length);
    # table: 2^w - 2 multiplications (fixed windows, the default), or one squaring and 2^(w-1) - 1
    # multiplications for the odd powers (DKG_SLIDING_WINDOW=1)
    w, nd = info["window_bits"], info["windows"]
    sliding = os.environ.get("DKG_SLIDING_WINDOW", "0") not in ("", "0")
    n_sqr = max(info["exponent_bits"] - 1, 0) + (1 if sliding else 0)
    n_mul = max(nd - 1, 0) + (((1 << (w - 1)) - 1) if sliding else ((1 << w) - 2)) + 2
    if info.get("pair_arithmetic"):
        K, M = info["pair_K"], info["pair_M"]
        blk, lo = K * K, K * (K + 1) // 2
        sqr = (M * (M + 1) // 2 + M * M) * blk + M * lo + 2 * M * M * blk + M * lo   # SQR(a) + 2ab
        mul = 3 * M * M * blk + M * lo + 2 * M * M * blk + M * lo                    # ad+bc, ac
    else:
        K, M = info["K"], info["M"]
        blk, lo = K * K, K * (K + 1) // 2
        sqr = (M * (M + 1) // 2 + M * M) * blk + M * lo
        mul = 2 * M * M * blk + M * lo
    return float(n_sqr * sqr + n_mul * mul)


def random_units(count: int, n_square: int, limbs: int, seed: int):
    """Uniform random residues below N^2 as limb rows (non-units have negligible probability)."""
    import numpy as np

    rng = np.random.default_rng(seed)
    arr = rng.integers(0, 2**32, size=(count, limbs), dtype=np.uint32)
    top_bits = n_square.bit_length() - 32 * (limbs - 1)
    # clear the top bit of N^2's width so every value is < 2^(bits-1) <= N^2
    arr[:, -1] &= np.uint32((1 << (top_bits - 1)) - 1)
    arr[:, 0] |= np.uint32(1)
    return arr


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self) -> None:
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu_index), "-lms", "200"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL,
            )
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        clocks, reasons, mx, power = [], set(), None, []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    clocks.append(float(f[1])); mx = float(f[2]); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if clocks:
            clocks.sort()
            out.update(sm_mhz=clocks[len(clocks) // 2], sm_max_mhz=mx, reasons=sorted(reasons),
                       samples=len(clocks), power_w_max=max(power) if power else None)
        return out


def cpu_baseline(dk, cores: int, sample: int, seed: int, cpython: bool = True) -> dict:
    """GMP mpz_powm (+ mpz_invert for the negative exponent) on `cores` pthreads over `sample`
    ciphertexts for each of the d+1 parties, plus the combination in CPython ints: the
    reference's CPU path with the [gmpy] extra (gmpy2.powmod wraps mpz_powm)."""
    from oracle import gmp

    dk = oracle_key(dk)
    n2 = dk.n * dk.n
    limbs = (n2.bit_length() + 31) // 32
    cts = random_units(sample, n2, limbs, seed)
    mod = gmp.int_to_limbs(n2, limbs)
    t0 = time.perf_counter()
    partial_rows = {}
    for pid in range(1, 2 * dk.t + 2):
        e = dk.keys[pid].partial_decrypt_exponent()
        el = gmp.int_to_limbs(abs(e), (abs(e).bit_length() + 31) // 32)
        out, _ = gmp.powm_batch_threads(cts, mod, el, e < 0, cores)
        partial_rows[pid] = gmp.limbs_to_ints(out)
    key1 = dk.keys[1]
    for i in range(sample):
        key1.decrypt({pid: partial_rows[pid][i] for pid in partial_rows})
    secs = time.perf_counter() - t0
    # secondary line: CPython pow on one core (the reference without its [gmpy] extra)
    out = {
        "value": sample / secs, "unit": UNIT, "cores": cores, "kind": "port", "secs": secs,
        "sample": f"{sample} ciphertexts x 3 parties GMP 6.3 mpz_powm/mpz_invert via oracle/c/gmp_batch.c on {cores} pthreads + CPython combine, {secs:.1f} s",
    }
    if cpython:
        k = min(4, sample)
        ints = gmp.limbs_to_ints(cts[:k])
        t1 = time.perf_counter()
        for pid in partial_rows:
            e = dk.keys[pid].partial_decrypt_exponent()
            for c in ints:
                got = pow(pow(c, -1, n2), -e, n2) if e < 0 else pow(c, e, n2)
            assert got == partial_rows[pid][k - 1], "GMP and CPython disagree"
        py_secs = time.perf_counter() - t1
        out["cpython_pow_1core"] = {"value": k / py_secs, "unit": UNIT,
                                    "sample": f"{k} ciphertexts x 3 parties, builtin pow"}
    return out


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dk = load_key()
    cores = os.cpu_count() or 1
    sample = args.ref_sample or max(cores * 96, 256)
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(dk, cores, max(cores, 16), 1, cpython=False)
    vals = []
    for s in range(args.steps):
        vals.append(cpu_baseline(dk, cores, sample, 100 + s, cpython=False))
    secs = sum(v["secs"] for v in vals)
    value = sample * args.steps / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": vals[-1]["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0, help="ciphertexts per GPU per step (0 = two full waves of the modexp kernel)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-sample", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch

    import protocols.distributed_keygen_b200 as eng
    from protocols.distributed_keygen_b200 import _native

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    if distributed:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    dk = load_key()
    shares = 2 * dk.t + 1
    keys = {}
    import math

    n_fac = math.factorial(dk.parties)
    for pid in range(1, shares + 1):
        share = eng.IntegerShares({pid: dk.shares[pid]}, 2 * dk.t, n_fac * n_fac, dk.parties)
        keys[pid] = eng.PaillierSharedKey(dk.n, dk.t, pid, share, dk.theta, device=local_rank)
    ctxs = {pid: key._modexp_ctx() for pid, key in keys.items()}
    comb = keys[1]._combine_ctx()
    L2, Ln = comb.n2_limbs, comb.n_limbs
    info = ctxs[1].info()
    B = args.batch or 2 * info["ctas"] * info["warps_per_cta"] * 32
    exps = {pid: keys[pid].partial_decrypt_exponent() for pid in keys}

    # ---- inputs: real encryptions are not needed for cost, but the result must be checkable:
    # use c = (1 + m N) * u^N style values?  r^N costs a modexp per element on the host, so take
    # uniformly random units and check partials bit-exactly against CPython pow on a sample, and
    # the combination on true encryptions in a small side batch.
    host_cts = random_units(B, dk.n * dk.n, L2, 1000 + rank)
    pinned_cts = torch.from_numpy(host_cts.view(np.int32)).pin_memory()
    d_cts = pinned_cts.cuda(non_blocking=True)
    d_partials = torch.empty((shares, B, L2), dtype=torch.int32, device="cuda")
    d_pstatus = torch.empty((shares, B), dtype=torch.uint8, device="cuda")
    d_plain = torch.empty((B, Ln), dtype=torch.int32, device="cuda")
    d_cstatus = torch.empty(B, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    modexp_events = []

    def device_step(record: bool) -> None:
        for s, pid in enumerate(range(1, shares + 1)):
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            ctxs[pid].modexp_device(d_cts.data_ptr(), d_partials[s].data_ptr(), d_pstatus[s].data_ptr(), B, stream)
            if record:
                e1.record()
                modexp_events.append((pid, e0, e1))
        comb.combine_device(d_partials.data_ptr(), d_plain.data_ptr(), d_cstatus.data_ptr(), B, stream)

    # ---- roofline denominator: measured on this GPU, now -------------------------------------
    import ctypes

    plain, carry = ctypes.c_double(0), ctypes.c_double(0)
    _native.check(_native.lib.dkg_measure_imad_peak(local_rank, ctypes.byref(plain), ctypes.byref(carry)))

    for _ in range(args.warmup):
        device_step(False)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        device_step(True)
    ev1.record()
    barrier()
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else {}
    elapsed_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())

    # ---- parity spot-check of what was just timed, against CPython integers --------------------
    # (the full parity suite against the oracle and the reference's golden vectors is tests/)
    from protocols.distributed_keygen_b200.limbs import limbs_to_ints

    part_host = d_partials.cpu().numpy().view(np.uint32)
    n2 = dk.n * dk.n
    for s, pid in enumerate(range(1, shares + 1)):
        e = exps[pid]
        for i in (0, B // 3, B - 1):
            c = limbs_to_ints(host_cts[i : i + 1])[0]
            want = pow(pow(c, -1, n2), -e, n2) if e < 0 else pow(c, e, n2)
            got = limbs_to_ints(part_host[s, i : i + 1])[0]
            assert got == want, f"rank {rank}: partial decryption mismatch party {pid} element {i}"
    assert int(d_pstatus.max().item()) == 0 and int(d_cstatus.max().item()) == 0
    plain_host = d_plain.cpu().numpy().view(np.uint32)
    theta_inv = pow(dk.theta, -1, dk.n)
    for i in (0, B // 2, B - 1):
        x = 1
        for s in range(shares):
            x = x * limbs_to_ints(part_host[s, i : i + 1])[0] % n2
        assert (x - 1) % dk.n == 0, "combined value minus one not divisible by N"
        assert limbs_to_ints(plain_host[i : i + 1])[0] == ((x - 1) // dk.n) * theta_inv % dk.n, "combine mismatch"

    # ---- end to end through the host-buffer API ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        pinned_partials = torch.empty((shares, B, L2), dtype=torch.int32).pin_memory()
        np_cts = pinned_cts.numpy().view(np.uint32)
        np_partials = pinned_partials.numpy().view(np.uint32)

        def e2e_step():
            for s, pid in enumerate(range(1, shares + 1)):
                out, status = keys[pid].partial_decrypt_limbs(np_cts)   # H2D + kernel + D2H
                np_partials[s] = out
            plain_out, cstatus = keys[1].decrypt_limbs(np_partials)     # H2D + kernel + D2H
            return plain_out, cstatus

        e2e_steps = max(1, min(args.steps, 3))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        secs = time.perf_counter() - t0
        t = torch.tensor([secs], dtype=torch.float64, device="cuda")
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_secs = float(t.item())
        h2d = shares * B * L2 * 4 + shares * B * L2 * 4
        d2h = shares * (B * L2 * 4 + B) + B * Ln * 4 + B
        e2e = {"value": world * B * e2e_steps / e2e_secs, "unit": UNIT, "steps": e2e_steps,
               "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world}

    # ---- true encryptions through the same kernels: decrypt(encrypt(m)) == m -------------------
    import random

    def encrypt_raw(n, m, r):  # (1 + m N) * r^N mod N^2, g = N + 1
        return (1 + m * n) * pow(r, n, n * n) % (n * n)

    rng = random.Random(7 + rank)
    ms = [rng.randrange(dk.n) for _ in range(64)]
    cts_small = [encrypt_raw(dk.n, m, rng.randrange(1, dk.n)) for m in ms]
    parts = {pid: keys[pid].partial_decrypt_batch(cts_small) for pid in keys}
    got = keys[1].decrypt_batch([{pid: parts[pid][i] for pid in keys} for i in range(64)])
    assert got == ms, "threshold decryption round trip failed"

    if rank == 0:
        # dominant kernel: modexp_fixed_kernel; per-launch duration from its own events
        durs = [e0.elapsed_time(e1) for (_, e0, e1) in modexp_events]
        avg_ms = sum(durs) / len(durs)
        ebits = sum(abs(exps[pid]).bit_length() for pid in exps) / len(exps)
        macs_per_launch = B * canonical_modexp_macs(int(round(ebits)), L2)
        achieved = macs_per_launch / (avg_ms * 1e-3) / 1e12
        actual = B * actual_modexp_macs(info) / (avg_ms * 1e-3) / 1e12
        # roofline denominator: the better of the two register-resident IMAD.WIDE probes measured
        # in this run (ptxas issues every IMAD.WIDE at 4-cycle intervals per sub-partition, so both
        # forms top out near 32 wide-MAC/clk/SM = 9.3 T/s at 1965 MHz)
        peak = max(plain.value, carry.value) / 1e12
        value = world * B * args.steps / (max_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (exact integer)",
            "data": "synthetic",
            "config": {
                "workload": WORKLOAD, "ciphertexts_per_gpu_per_step": B, "parties": dk.parties,
                "threshold": dk.t, "modulus_bits": dk.n.bit_length(), "n2_limbs": L2,
                "exponent_bits": [abs(exps[p]).bit_length() * (1 if exps[p] > 0 else -1) for p in sorted(exps)],
                "kernel_shape": info, "parallelism": f"index-sharded x{world}, no collective",
                "cache": "inputs+outputs per step (%.0f MB) larger than L2" % ((shares + 1) * B * L2 * 4 / 1e6),
            },
            "roofline": {
                "bound": "imad", "achieved": achieved, "peak": peak, "unit": "T wide-MAC/s",
                "frac": achieved / peak,
                # DRAM bytes of one launch of this kernel: ncu dram__bytes_read.sum + dram__bytes_write.sum
                # captured once on a 113 664-ciphertext launch (profiles/r01_ncu_nsq_traffic_final.csv:
                # 37.80 GB + 8.72 GB, window tables spilling out of L2), scaled to this launch's size;
                # only known for the pair-arithmetic kernel at 2048-bit N
                "traffic": (46.523e9 * B / 113664.0) if (info.get("pair_arithmetic") and info.get("pair_K") == 14) else None,
                "traffic_unit": "bytes per launch (DRAM, from the committed ncu capture)",
                # the HBM side of the roofline, to show the kernel is nowhere near it
                "hbm": hbm_side(46.523e9 * B / 113664.0, avg_ms) if (info.get("pair_arithmetic") and info.get("pair_K") == 14) else None,
                "kernel": ("modexp_nsq_kernel<%d,%d>" % (info["pair_K"], info["pair_M"])) if info.get("pair_arithmetic")
                else ("modexp_fixed_kernel<%d,%d>" % (info["K"], info["M"])),
                "actual_wide_mac": actual, "frac_actual": actual / peak,
                "avg_launch_ms": avg_ms, "algorithmic_macs_per_launch": macs_per_launch,
                "peak_source": "dkg_measure_imad_peak, measured in this run: max of register-resident mad.wide.u32 (plain) and mad.lo.cc/madc.hi.cc (carry chain) probes",
                "peak_plain_mad_wide": plain.value / 1e12,
                "peak_carry_chain": carry.value / 1e12,
                "note": "achieved = canonical work of SURVEY 8(d) (2L^2+L per modmul at L=limbs of N^2, squarings as multiplies); "
                        "the kernel computes the same results with fewer multiplies (block squaring; for N^2 moduli pair arithmetic "
                        "modulo N, csrc/dkg_nsq.cuh), so achieved/peak can exceed 1; frac_actual = multiplies really executed / peak",
            },
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if e2e is not None:
            line["e2e"] = e2e
        cores = os.cpu_count() or 1
        if world == 1:
            line["cpu_baseline"] = cpu_baseline(dk, cores, args.cpu_sample or max(cores * 96, 256), 5)
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
