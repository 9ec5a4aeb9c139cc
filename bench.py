#!/usr/bin/env python
"""bench.py -- prover throughput of the SHA-256 ZK proof (BASELINE.json configs[1],
BM_ShaZK_fp2_128/1: 1-block flatsha256 circuit, Ligero over GF(2^128), rate 7,
132 queries) on N B200s, next to the reference CPU prover on the box's host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

A step = one batch of B independent proofs (fresh witness buffer, RNG bytes and
transcript per proof) through the whole prover hot path: tableau layout, RS row
encode, Merkle commit, on-device Fiat-Shamir transcript, eval_circuit, the layered
sumcheck, Ligero prove and proof serialization.  `value` = proofs/s with inputs
and outputs resident in HBM: the K steps are issued round robin on three
contexts/streams (so one batch's thinly parallel kernels run under another's
sumcheck) and timed by one CUDA-event interval that encloses all of them, max
over ranks; `streams.one_stream` is the same K steps back to back on one stream.
`e2e` = the same through the host-pointer C-ABI call (lf_zk_prove_batch) with
pinned HOST buffers, H2D and D2H inside the timed region, three batches in flight
from three host threads; `e2e.one_batch_at_a_time` is the strictly serial figure.
Independent proofs shard across GPUs with no collective (weak scaling: B proofs
per GPU).
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "sha256_zk_prover_throughput"
UNIT = "proofs/s"
WORKLOAD = "BM_ShaZK_fp2_128/1: 1-block SHA-256 ZK proof, GF(2^128), rate 7, nreq 132"


WORKLOADS = {
    # name -> (fixture, field id, description)
    "sha": ("sha1_gf128", 4, WORKLOAD),
    "ecdsa": ("ecdsa1_p256", 1, "BM_ECDSAZKProver/1: ECDSA P-256 verification ZK proof, Fp256, rate 7, nreq 132"),
}


def load_fixture(name="sha"):
    from fixtures import load
    return load(WORKLOADS[name][0])


def witness_rows(name, B):
    """B witnesses cycling through the distinct fixtures of the workload (other messages / signatures)"""
    import numpy as np
    from fixtures import load_witnesses
    W = load_witnesses(WORKLOADS[name][0])
    return np.ascontiguousarray(W[np.arange(B) % W.shape[0]]), W.shape[0]


def ref_build_info():
    """compiler and flags the reference arm (oracle/_ref) was built with (written by __graft_entry__.build)"""
    try:
        return json.load(open(os.path.join(ROOT, "oracle", "_ref", "build_info.json")))
    except Exception:
        return dict(compiler="g++ (version not recorded)", flags="see oracle/ref_build/Makefile")


def rng_stream(seed, n):
    import numpy as np
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8)


# ----------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified reference prover (oracle/_ref) on
# the host cores, N independent single-threaded provers (the library has no
# threads of its own, docs/content/en/docs/benchmarks.md:7).
# ----------------------------------------------------------------------------
def cpu_reference_throughput(nthreads, per_thread, workload="sha"):
    from oracle import refapi, portapi
    circ, wit = load_fixture(workload)
    fid = WORKLOADS[workload][1]
    rng = rng_stream(1, 1 << 19)
    if refapi.available():
        c = refapi.Circuit(fid, circ)
        secs, lat = c.bench(wit, rng, nthreads=nthreads, per_thread=per_thread)
        kind = "reference"
    else:  # the reference could not be built here: fall back to the oracle port
        c = portapi.Circuit(fid, circ)
        t0 = time.time()
        for _ in range(per_thread):
            c.prove(wit, rng)
        secs, lat, nthreads = time.time() - t0, [], 1
        kind = "port"
    n = nthreads * per_thread
    med = sorted(lat)[len(lat) // 2] if lat else secs / n * 1e3
    return dict(value=n / secs, unit=UNIT, cores=nthreads, kind=kind,
                sample=f"{nthreads} threads x {per_thread} proofs of the same workload ({secs:.2f} s wall)",
                ms_per_proof_1thread=med, build=ref_build_info())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nthreads = os.cpu_count() or 1
    per_thread = 40
    vals, secs_all = [], []
    for i in range(args.warmup + args.steps):
        t0 = time.time()
        r = cpu_reference_throughput(nthreads, per_thread)
        if i >= args.warmup:
            vals.append(r["value"])
            secs_all.append(time.time() - t0)
    v = sum(vals) / len(vals)
    r["value"] = v
    line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * sum(secs_all) / len(secs_all), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="gf2^128 (u32 limbs)", data="synthetic", impl="reference",
                config=dict(workload=WORKLOAD, proofs_per_step=nthreads * per_thread, host_threads=nthreads),
                cpu_baseline=r,
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


# ----------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.p = [], None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v == "Active"})
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


def stage_rates(info, stage_ms, B, mul_peak, sha_peak):
    """per-stage achieved integer-op rates against the measured peaks of this box (lf_microbench):
    field multiplications/s of the stages that are multiply bound, compressions/s of the Merkle commit"""
    out = {}
    for stage, key in (("rs_encode", "rs_mults"), ("eval_circuit", "eval_mults"), ("sumcheck", "sumcheck_mults"),
                       ("ligero_prove", "ligero_mults")):
        ms = stage_ms.get(stage, 0.0)
        if ms > 0:
            g = info[key] * B / (ms * 1e-3) / 1e9
            out[stage] = dict(ms=ms, mults_per_proof=info[key], achieved_gmul_s=g, frac_of_mul_peak=g / mul_peak)
    ms = stage_ms.get("merkle", 0.0)
    if ms > 0:
        g = info["merkle_compressions"] * B / (ms * 1e-3) / 1e9
        out["merkle"] = dict(ms=ms, compressions_per_proof=info["merkle_compressions"], achieved_gcomp_s=g,
                             frac_of_sha_peak=g / sha_peak)
    return out


def measure_other(lf, ctx, stream, workload, B, steps=3, reduce_max=None, world=1):
    """device-resident and end-to-end proofs/s of another circuit; under torchrun every rank runs it on its own
    GPU and the times are max-reduced (reduce_max), so the value is the whole-job aggregate"""
    import numpy as np
    import torch
    fixture, fid, desc = WORKLOADS[workload]
    circ, wit = load_fixture(workload)
    circuit = lf.Circuit(ctx, fid, circ)
    prover = lf.ZkProver(circuit)
    info = circuit.info
    wb, rb, pb = info["witness_bytes"], info["rng_bytes"], info["max_proof_bytes"]
    # room for a few redrawn samples per proof (prime fields: Field::sample rejects a draw >= p, 2^-32 each)
    rstride = (rb + 8 * info["rng_redraw_bytes"] + 15) & ~15
    gen = torch.Generator().manual_seed(77)
    W, ndistinct = witness_rows(workload, B)
    h_wit = torch.from_numpy(W).pin_memory()
    h_rng = torch.randint(0, 256, (B, rstride), dtype=torch.uint8, generator=gen).pin_memory()
    h_out = torch.empty((B, pb), dtype=torch.uint8).pin_memory()
    h_len = torch.zeros(B, dtype=torch.int64).pin_memory()
    h_st = torch.zeros(B, dtype=torch.int32).pin_memory()
    d_wit, d_rng = h_wit.cuda(), h_rng.cuda()
    d_out = torch.empty((B, pb), dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(B, dtype=torch.int64, device="cuda")
    d_st = torch.zeros(B, dtype=torch.int32, device="cuda")
    rmax = reduce_max or (lambda x: x)

    def dev():
        prover.prove_batch_ptr(B, d_wit.data_ptr(), d_rng.data_ptr(), rstride, d_out.data_ptr(), pb,
                               d_len.data_ptr(), d_st.data_ptr(), device=True)

    def host():
        prover.prove_batch_ptr(B, h_wit.data_ptr(), h_rng.data_ptr(), rstride, h_out.data_ptr(), pb,
                               h_len.data_ptr(), h_st.data_ptr(), device=False)
    for _ in range(3):
        dev()
    torch.cuda.synchronize()
    assert int(d_st.abs().sum().item()) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rmax(0.0)  # rendezvous
    e0.record(stream)
    for _ in range(steps):
        dev()
    e1.record(stream)
    torch.cuda.synchronize()
    ms_one = rmax(e0.elapsed_time(e1))
    # the same with three batches in flight on three contexts / streams (as the headline workload is run)
    NS = 3
    xs = [torch.cuda.Stream() for _ in range(NS - 1)]
    xc = [lf.Context(torch.cuda.current_device(), stream=s_.cuda_stream) for s_ in xs]
    xp = [lf.ZkProver(lf.Circuit(c_, fid, circ)) for c_ in xc]
    xo = [(torch.empty((B, pb), dtype=torch.uint8, device="cuda"), torch.zeros(B, dtype=torch.int64, device="cuda"),
           torch.zeros(B, dtype=torch.int32, device="cuda")) for _ in range(NS - 1)]

    def dev_x(k):
        xp[k].prove_batch_ptr(B, d_wit.data_ptr(), d_rng.data_ptr(), rstride, xo[k][0].data_ptr(), pb,
                              xo[k][1].data_ptr(), xo[k][2].data_ptr(), device=True)
    for k in range(NS - 1):
        dev_x(k)
        dev_x(k)
    torch.cuda.synchronize()
    nsteps3 = NS * max(2, steps // 1)
    rmax(0.0)
    e0.record(stream)
    for s_ in xs:
        s_.wait_event(e0)
    for i in range(nsteps3):
        if i % NS == 0:
            dev()
        else:
            dev_x(i % NS - 1)
    for s_ in xs:
        ej = torch.cuda.Event()
        ej.record(s_)
        stream.wait_event(ej)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = rmax(e0.elapsed_time(e1)) * steps / nsteps3   # per `steps` steps, comparable with ms_one
    for o_ in xo:
        assert int(o_[2].abs().sum().item()) == 0
    # end to end with the same three batches in flight: one host thread, context, stream and pinned
    # buffers each, so that one batch's PCIe copies overlap the other batches' kernels (as the headline's e2e)
    hx = [[h_wit.clone().pin_memory(), h_rng.clone().pin_memory(), torch.empty((B, pb), dtype=torch.uint8).pin_memory(),
           torch.zeros(B, dtype=torch.int64).pin_memory(), torch.zeros(B, dtype=torch.int32).pin_memory()]
          for _ in range(NS - 1)]

    def host_x(k):
        h = hx[k]
        xp[k].prove_batch_ptr(B, h[0].data_ptr(), h[1].data_ptr(), rstride, h[2].data_ptr(), pb, h[3].data_ptr(),
                              h[4].data_ptr(), device=False)
    host()
    for k in range(NS - 1):
        host_x(k)
    n_pipe = 2
    rmax(0.0)
    t0 = time.perf_counter()
    ths = [threading.Thread(target=lambda k=k: [host_x(k) for _ in range(n_pipe)]) for k in range(NS - 1)]
    for th in ths:
        th.start()
    for _ in range(n_pipe):
        host()
    for th in ths:
        th.join()
    t_pipe = rmax(time.perf_counter() - t0)
    for h in hx:
        assert int(h[4].abs().sum().item()) == 0
    assert int(h_st.abs().sum().item()) == 0
    e2e_pipe = world * B * NS * n_pipe / t_pipe
    del xp, xo, hx
    for c_ in xc:
        c_.close()
    prover.set_profiling(True)
    dev()
    stages = prover.stage_ms()
    kernel_ms = prover.kernel_ms()
    prover.set_profiling(False)
    # single-proof latency (batch of one, device resident)
    def one():
        prover.prove_batch_ptr(1, d_wit.data_ptr(), d_rng.data_ptr(), rstride, d_out.data_ptr(), pb,
                               d_len.data_ptr(), d_st.data_ptr(), device=True)
    for _ in range(2):
        one()
    e0.record(stream)
    for _ in range(5):
        one()
    e1.record(stream)
    torch.cuda.synchronize()
    lat1 = e0.elapsed_time(e1) / 5
    host()
    rmax(0.0)
    t0 = time.perf_counter()
    for _ in range(2):
        host()
    t1 = rmax(time.perf_counter() - t0)
    assert int(h_st.abs().sum().item()) == 0
    # the verifier on the proofs just made (host buffers in, status out): BM_ECDSAZKVerifier's counterpart
    npubb = info["npub_in"] * info["kbytes"]
    proofs = [bytes(h_out[i, :int(h_len[i])].numpy()) for i in range(B)]
    pubs = np.ascontiguousarray(W[:, :npubb]) if npubb else None
    ver = lf.ZkVerifier(circuit)
    st, _ = ver.verify_batch(pubs, proofs)
    assert (st == 0).all(), "the GPU verifier rejected a GPU proof"
    # raw-pointer timing (no Python packing in the timed region)
    import ctypes as C
    from longfellow_zk_b200 import _native
    lens64 = np.array([len(p_) for p_ in proofs], np.uint64)
    vst, vwhy = np.zeros(B, np.int32), np.zeros(B, np.int32)
    pp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    hout_np = h_out.numpy()

    def verify_raw():
        _native.check(_native.lib().lf_zk_verify_batch(circuit._h, B, pp(pubs), pp(hout_np), pb, pp(lens64), b"test", 4,
                                                       pp(vst), pp(vwhy)))
    verify_raw()
    rmax(0.0)
    t0 = time.perf_counter()
    for _ in range(2):
        verify_raw()
    tv = rmax(time.perf_counter() - t0)
    assert (vst == 0).all()
    return dict(workload=desc, proofs_per_step=B, distinct_witnesses=ndistinct,
                value=world * B * steps / (ms * 1e-3), unit=UNIT,
                streams=dict(n=3, one_stream=dict(value=world * B * steps / (ms_one * 1e-3))),
                e2e=dict(value=e2e_pipe, unit=UNIT, how="lf_zk_prove_batch with pinned host buffers, three batches in "
                         "flight (one host thread, context and stream each)",
                         one_batch_at_a_time=dict(value=world * 2 * B / t1)), stage_ms=stages,
                kernel_ms={k: dict(ms=v[0], launches=v[1]) for k, v in kernel_ms.items()},
                latency_ms_per_proof_batch1=lat1,
                kernels=stage_rates(info, stages, B, ctx.microbench(4 if fid == 1 else 2), ctx.microbench(3)),
                verifier=dict(value=world * 2 * B / tv, unit="proofs verified/s", ms_per_proof=1e3 * tv / (2 * B),
                              how="lf_zk_verify_batch, host buffers in / status out, batch of %d" % B),
                proof_bytes=int(d_len[0].item()), total_mults_per_proof=info["total_mults"])


def cpu_verify_ms(workload, reps=5):
    """the reference's ZkProof::read + ZkVerifier on one host thread: ms per proof (median)"""
    from oracle import refapi
    import numpy as np
    if not refapi.available():
        return None
    circ, wit = load_fixture(workload)
    fid = WORKLOADS[workload][1]
    c = refapi.Circuit(fid, circ)
    info = refapi.circuit_info(fid, circ)
    kb = 16 if fid == 4 else 32
    pr = c.prove(wit, rng_stream(1, 1 << 19))["proof"]
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        assert c.verify(wit[:info["npub_in"] * kb], pr) == 0
        ts.append(1e3 * (time.perf_counter() - t0))
    return sorted(ts)[len(ts) // 2]


FIELDS_CONFIG1 = [("fp256_p256_fp2", 1), ("bn254_fp4", 100), ("fp128", 101), ("goldilocks_f64", 102)]


# the other instance sizes the reference publishes (Apple M4, one thread, ms per proof)
PUBLISHED_M4_MS = {"sha2_gf128": 9.60, "sha4_gf128": 18.73, "sha8_gf128": 35.39, "sha16_gf128": 65.62,
                   "sha32_gf128": 125.23, "sha33_gf128": 132.71, "ecdsa2_p256": 26.51, "ecdsa3_p256": 38.32}


def measure_sizes(lf, ctx, stream, batch=64):
    """BM_ShaZK_fp2_128/2..33 and BM_ECDSAZKProver/2,3: one proof (latency) and a batch of `batch` proofs,
    device resident, with the unmodified reference on one host thread of this box and the published M4 figure"""
    import numpy as np
    import torch
    from fixtures import SIZE_NAMES, load_size
    from oracle import refapi
    out = {}
    for name in SIZE_NAMES:
        try:
            circ, wit, g = load_size(name)
            fid = g["field_id"]
            c = lf.Circuit(ctx, fid, circ)
            p = lf.ZkProver(c)
            info = c.info
            pb = info["max_proof_bytes"]
            rstride = (info["rng_bytes"] + 8 * info["rng_redraw_bytes"] + 15) & ~15
            B = batch
            d_wit = torch.from_numpy(np.frombuffer(wit, np.uint8).copy()).repeat(B, 1).cuda()
            d_rng = torch.randint(0, 256, (B, rstride), dtype=torch.uint8,
                                  generator=torch.Generator().manual_seed(5)).cuda()
            d_out = torch.empty((B, pb), dtype=torch.uint8, device="cuda")
            d_len = torch.zeros(B, dtype=torch.int64, device="cuda")
            d_st = torch.zeros(B, dtype=torch.int32, device="cuda")

            def run(n):
                p.prove_batch_ptr(n, d_wit.data_ptr(), d_rng.data_ptr(), rstride, d_out.data_ptr(), pb,
                                  d_len.data_ptr(), d_st.data_ptr(), device=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            res = {}
            for n, reps in ((1, 3), (B, 2)):
                run(n)
                torch.cuda.synchronize()
                e0.record(stream)
                for _ in range(reps):
                    run(n)
                e1.record(stream)
                torch.cuda.synchronize()
                assert int(d_st[:n].abs().sum().item()) == 0
                res[n] = e0.elapsed_time(e1) / reps
            rec = dict(nterms=info["nterms"], tableau=[info["nrow"], info["block_enc"]], proof_bytes=int(d_len[0].item()),
                       latency_ms_per_proof_batch1=res[1], batch=B, ms_per_proof_in_batch=res[B] / B,
                       proofs_per_s_in_batch=1e3 * B / res[B], published_m4_1thread_ms=PUBLISHED_M4_MS.get(name))
            if refapi.available():
                rng = rng_stream(3, 1 << 22)
                _, lat = refapi.Circuit(fid, circ).bench(wit, rng, nthreads=1, per_thread=2)
                rec["reference_1thread_ms_this_host"] = min(lat)
            out[name] = rec
            del p
            c.close()
        except Exception as ex:
            out[name] = dict(error=repr(ex))
    return out


def measure_config1(ctx):
    """BASELINE config 1: FFT (n = 2^16) and Reed-Solomon encode per field, a single transform and 1024 rows at once,
    device resident (CUDA events on the context stream), next to FFT<Field>::fftb / ReedSolomon::interpolate of
    the unmodified reference on one host thread."""
    from oracle import refapi
    have_ref = refapi.available()
    out = {}
    for name, fid in FIELDS_CONFIG1:
        r = dict(fft_65536_ms=ctx.fft_time_ms(fid, 65536, reps=20),
                 fft_65536_x1024rows_ms=ctx.fft_time_rows_ms(fid, 65536, 1024 if fid != 1 else 512, reps=3),
                 fft_rows=1024 if fid != 1 else 512,
                 rs_65536_262144_ms=ctx.rs_time_ms(fid, 65536, 262144, 1, reps=3),
                 rs_455_4096_x1024rows_ms=ctx.rs_time_ms(fid, 455, 4096, 1024, reps=5))
        if have_ref:
            r["cpu_1thread"] = dict(fft_65536_ms=1e3 * refapi.fft_bench(fid, 65536, 2),
                                    rs_65536_262144_ms=1e3 * refapi.rs_bench(fid, 65536, 262144, 1),
                                    rs_455_4096_ms_per_row=1e3 * refapi.rs_bench(fid, 455, 4096, 3))
        out[name] = r
    r = dict(rs_16384_65536_ms=ctx.rs_time_ms(4, 16384, 65536, 1, reps=5),
             rs_16384_65536_x64rows_ms=ctx.rs_time_ms(4, 16384, 65536, 64, reps=3),
             rs_455_4096_x1024rows_ms=ctx.rs_time_ms(4, 455, 4096, 1024, reps=5),
             rs_455_4096_x20480rows_ms=ctx.rs_time_ms(4, 455, 4096, 20480, reps=5))
    if have_ref:
        r["cpu_1thread"] = dict(rs_16384_65536_ms=1e3 * refapi.rs_bench(4, 16384, 65536, 3),
                                rs_455_4096_ms_per_row=1e3 * refapi.rs_bench(4, 455, 4096, 10))
    out["gf2_128_lch14"] = r
    return out


def run_ours(args):
    import numpy as np
    import torch
    import longfellow_zk_b200 as lf

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the prover has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    # stdout is the one JSON line: whatever a library prints there (NCCL's version / INFO lines under NCCL_DEBUG,
    # which stays as the driver set it) is sent to stderr at the file-descriptor level
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    circ, wit = load_fixture()
    # the prover launches on an explicit (non-default) stream; the CUDA events
    # below are recorded on that same stream
    stream = torch.cuda.Stream()
    assert stream.cuda_stream != 0
    ctx = lf.Context(local, stream=stream.cuda_stream)
    circuit = lf.Circuit(ctx, lf.FIELD_GF2_128, circ)
    prover = lf.ZkProver(circuit)
    info = circuit.info
    B = args.batch
    wb, rb, pb = info["witness_bytes"], info["rng_bytes"], info["max_proof_bytes"]
    rstride = (rb + 15) & ~15

    # synthetic inputs: the proofs of a batch cycle through distinct witnesses (different SHA-256 messages),
    # every proof has its own RNG stream
    gen = torch.Generator().manual_seed(1234 + rank)
    W, ndistinct = witness_rows("sha", B)
    h_wit = torch.from_numpy(W).pin_memory()
    h_rng = torch.randint(0, 256, (B, rstride), dtype=torch.uint8, generator=gen).pin_memory()
    h_out = torch.empty((B, pb), dtype=torch.uint8).pin_memory()
    h_len = torch.zeros(B, dtype=torch.int64).pin_memory()
    h_st = torch.zeros(B, dtype=torch.int32).pin_memory()
    d_wit, d_rng = h_wit.cuda(), h_rng.cuda()
    d_out = torch.empty((B, pb), dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(B, dtype=torch.int64, device="cuda")
    d_st = torch.zeros(B, dtype=torch.int32, device="cuda")

    def step_dev():
        prover.prove_batch_ptr(B, d_wit.data_ptr(), d_rng.data_ptr(), rstride, d_out.data_ptr(), pb,
                               d_len.data_ptr(), d_st.data_ptr(), device=True)

    # More contexts on their own streams: consecutive steps (batches) go round robin over the
    # streams, so the thinly parallel kernels of one batch (zero-block hashing of the transcript,
    # proof serialisation, Merkle tree top, the small sumcheck rounds) run under another batch's large kernels.
    NS = 3  # streams in flight (tools/streams_sweep.py)
    xstreams = [torch.cuda.Stream() for _ in range(NS - 1)]
    xctx = [lf.Context(local, stream=s.cuda_stream) for s in xstreams]
    xprover = [lf.ZkProver(lf.Circuit(cx, lf.FIELD_GF2_128, circ)) for cx in xctx]
    xout = [(torch.empty((B, pb), dtype=torch.uint8, device="cuda"), torch.zeros(B, dtype=torch.int64, device="cuda"),
             torch.zeros(B, dtype=torch.int32, device="cuda")) for _ in range(NS - 1)]

    def step_dev_x(k):
        o = xout[k]
        xprover[k].prove_batch_ptr(B, d_wit.data_ptr(), d_rng.data_ptr(), rstride, o[0].data_ptr(), pb,
                                   o[1].data_ptr(), o[2].data_ptr(), device=True)

    def reduce_max(x):
        """max over ranks of a host scalar (also the rendezvous before a timed region)"""
        torch.cuda.synchronize()
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_streams(steps):
        """K steps issued round robin on the NS streams; one CUDA-event interval on `stream`
        that starts before the first launch on any stream and ends after the last on all."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s in xstreams:
            s.wait_event(e0)
        for i in range(steps):
            if i % NS == 0:
                step_dev()
            else:
                step_dev_x(i % NS - 1)
        for s in xstreams:
            ej = torch.cuda.Event()
            ej.record(s)
            stream.wait_event(ej)
        e1.record(stream)
        torch.cuda.synchronize()
        t = reduce_max(e0.elapsed_time(e1))
        barrier()
        return t

    def step_host():
        prover.prove_batch_ptr(B, h_wit.data_ptr(), h_rng.data_ptr(), rstride, h_out.data_ptr(), pb,
                               h_len.data_ptr(), h_st.data_ptr(), device=False)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, use_events):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        t = reduce_max(e0.elapsed_time(e1) if use_events else (t1 - t0) * 1e3)
        barrier()
        return t

    # ---- warm-up + correctness of what is being timed
    for _ in range(max(args.warmup, 3)):
        step_dev()
        for k in range(NS - 1):
            step_dev_x(k)
    torch.cuda.synchronize()
    assert int(d_st.abs().sum().item()) == 0, "prover reported failures"
    lens = d_len.cpu().numpy()
    assert (lens > 100000).all()
    for o in xout:  # same inputs, same proofs on every stream
        assert int(o[2].abs().sum().item()) == 0 and torch.equal(d_out[:4], o[0][:4])

    # ---- device-resident throughput (value): K steps, round robin over the NS streams
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ctx.launch_count + sum(cx.launch_count for cx in xctx)
    ms_total = timed_streams(args.steps)
    launches = ctx.launch_count + sum(cx.launch_count for cx in xctx) - l0
    clocks = sampler.stop() if sampler else None
    value = world * B * args.steps / (ms_total * 1e-3)
    # the same K steps back to back on ONE stream (no overlap between batches)
    ms_one_stream = timed(step_dev, args.steps, use_events=True)

    # ---- per-stage and per-kernel-class device times of three more batches (roofline of the dominant kernel)
    prover.set_profiling(True)
    stage_acc, kacc = {}, {}
    nprof = 3
    for _ in range(nprof):
        step_dev()
        for k, v in prover.stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v / nprof
        for k, (ms_k, n_k) in prover.kernel_ms().items():
            a = kacc.setdefault(k, [0.0, 0])
            a[0] += ms_k / nprof
            a[1] = n_k
    prover.set_profiling(False)

    # ---- end to end through the host-pointer C-ABI call (pinned host buffers)
    for _ in range(2):
        step_host()
    e2e_steps = max(2, min(args.steps, 5))
    ms_e2e = timed(step_host, e2e_steps, use_events=False)
    e2e_serial = world * B * e2e_steps / (ms_e2e * 1e-3)
    assert int(h_st.abs().sum().item()) == 0
    # the same call from NS host threads, each with its own context / stream / pinned buffers, so
    # that one batch's PCIe copies overlap the other batches' kernels (what a serving loop does)
    hx = [[h_wit.clone().pin_memory(), h_rng.clone().pin_memory(), torch.empty((B, pb), dtype=torch.uint8).pin_memory(),
           torch.zeros(B, dtype=torch.int64).pin_memory(), torch.zeros(B, dtype=torch.int32).pin_memory()]
          for _ in range(NS - 1)]

    def step_host_x(k):
        h = hx[k]
        xprover[k].prove_batch_ptr(B, h[0].data_ptr(), h[1].data_ptr(), rstride, h[2].data_ptr(), pb,
                                   h[3].data_ptr(), h[4].data_ptr(), device=False)
    for k in range(NS - 1):
        step_host_x(k)

    def all_threads():
        ths = [threading.Thread(target=lambda k=k: [step_host_x(k) for _ in range(e2e_steps)]) for k in range(NS - 1)]
        for th in ths:
            th.start()
        for _ in range(e2e_steps):
            step_host()
        for th in ths:
            th.join()
    ms_pipe = timed(all_threads, 1, use_events=False)
    e2e_value = world * B * NS * e2e_steps / (ms_pipe * 1e-3)
    for h in hx:
        assert int(h[4].abs().sum().item()) == 0 and int(h[3][0].item()) == int(h_len[0].item())
    ms_e2e_step = ms_pipe / (NS * e2e_steps)
    # the proofs that came back through the host path equal the device-resident ones
    torch.cuda.synchronize()
    n0 = int(h_len[0].item())
    assert n0 == int(lens[0]) and bytes(h_out[0, :n0].numpy()) == bytes(d_out[0, :n0].cpu().numpy())

    # ---- single-proof latency (batch of one, device resident)
    d_len1 = torch.zeros(1, dtype=torch.int64, device="cuda")
    d_st1 = torch.zeros(1, dtype=torch.int32, device="cuda")

    def step_one():
        prover.prove_batch_ptr(1, d_wit.data_ptr(), d_rng.data_ptr(), rstride, d_out.data_ptr(), pb,
                               d_len1.data_ptr(), d_st1.data_ptr(), device=True)
    for _ in range(3):
        step_one()
    lat_ms = timed(step_one, 10, use_events=True) / 10

    # ---- BASELINE config 5 names SHA-256 AND ECDSA at N GPUs, config 3 the mdoc proof: every rank runs them
    other = {}
    skip = bool(args.headline_only)   # same on every rank
    try:
        if skip:
            other["ecdsa_p256"] = dict(skipped="--headline-only")
        else:
            other["ecdsa_p256"] = measure_other(lf, ctx, stream, "ecdsa", min(B, 1024), reduce_max=reduce_max,
                                                world=world)
    except Exception as ex:  # the headline line must not depend on the extra workload
        other["ecdsa_p256"] = dict(error=repr(ex))
        reduce_max(0.0)
    try:
        if skip:
            raise RuntimeError("skipped: --headline-only")
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import mdoc_bench
        md = mdoc_bench.measure(batches=(1, 128), reps=2, device=local) if rank == 0 else None
        barrier()
        md2 = mdoc_bench.measure_two_in_flight(B=128, rounds=2, device=local, rendezvous=barrier)
        secs = reduce_max(md2["seconds"])
        if rank == 0:
            nproofs = world * md2["batches"] * 128
            other["mdoc"] = dict(workload="BM_MdocProver: kZkSpecs[0], mdoc_tests[0] + age_over_18; hash circuit GF(2^128) "
                                          "7.76 M terms 266x4151, signature circuit Fp256 482 k terms 19x4096; "
                                          "commit+commit+prove+prove through the host-pointer C ABI (H2D/D2H included)",
                                 latency_ms_per_proof_batch1=md["batches"][0]["ms_total"],
                                 value=nproofs / secs, unit=UNIT, proofs_per_step=128, n_gpus=world,
                                 ms_per_proof=1e3 * secs / nproofs,
                                 how="per GPU two batches of 128 in flight (two host threads, contexts and circuit objects)",
                                 one_batch_at_a_time=dict(value=md["batches"][1]["proofs_per_s"],
                                                          ms_per_proof=md["batches"][1]["ms_per_proof"]),
                                 detail=md)
    except Exception as ex:
        other["mdoc"] = dict(error=repr(ex))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel.  The path is integer-pipe bound (GF(2^128) products out of IMAD.WIDE
    # and LOP3), so the bound that is reported is the integer one; HBM is the secondary figure.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "6650 GB/s (fallback)"
    top = max(stage_acc, key=stage_acc.get)
    sc_ms = stage_acc["sumcheck"]
    gmul_peak = ctx.microbench(2)   # measured GF(2^128) multiply rate of this formulation (Gmul/s)
    imad_gops = ctx.microbench(0)   # IMAD.WIDE, G warp-lane ops/s
    lop3_gops = ctx.microbench(1)
    # one product = 144 IMAD.WIDE + 251 LOP3 (cuobjdump of gf128_mul): the pipe-level ceiling of the formulation
    pipe_bound = min(imad_gops / 144.0, lop3_gops / 251.0)
    kname = max(kacc, key=lambda k: kacc[k][0]) if kacc else "k_zk_sumcheck"
    kms, kn = kacc.get(kname, (sc_ms, 1))
    if kname == "k_sc_eval":
        k_bytes, k_mults = info["flat_eval_alg_bytes"] * B, info["flat_eval_mults"] * B
    elif kname == "k_sc_bind":
        k_bytes, k_mults = info["flat_bind_alg_bytes"] * B, info["flat_bind_mults"] * B
    else:
        k_bytes, k_mults = info["sumcheck_alg_bytes"] * B, info["sumcheck_mults"] * B
    k_gmul = k_mults / (kms * 1e-3) / 1e9
    k_gbs = k_bytes / (kms * 1e-3) / 1e9
    # DRAM bytes of one launch of that kernel from the committed `ncu --set full` capture of the launched variant
    traffic, traffic_src = None, None
    try:
        import re
        cap_file = "profiles/r2_sc_eval_final.txt"
        cap = open(os.path.join(ROOT, cap_file)).read().split("kernel:")[1]
        rd = float(re.search(r"dram__bytes_read.sum\s+([\d.]+)\s+Mbyte", cap).group(1))
        wr = float(re.search(r"dram__bytes_write.sum\s+([\d.]+)\s+Mbyte", cap).group(1))
        traffic = (rd + wr) * 1e6
        traffic_src = cap_file + ": dram__bytes_read.sum + dram__bytes_write.sum of ONE k_sc_eval launch (layer 11 " \
            "round 4 of 1024 proofs: 19351 entries + 5700 wires per proof, 886 MB algorithmic; the L2 holds part of it)"
    except Exception:
        pass
    roofline = dict(bound="integer", kernel=kname, achieved=k_gmul, peak=gmul_peak, unit="Gmul/s (GF(2^128) products)",
                    frac=k_gmul / gmul_peak,
                    peak_source="lf_microbench(2): the same multiply formulation in a register-resident loop, measured in "
                                "this run",
                    launches_per_step=kn, kernel_ms_per_launch=kms / max(kn, 1), kernel_ms_per_step=kms,
                    share_of_step=kms / sum(stage_acc.values()),
                    algorithmic_mults_per_launch=k_mults / max(kn, 1),
                    algorithmic_bytes_per_launch=k_bytes / max(kn, 1),
                    traffic=traffic, traffic_source=traffic_src,
                    pipes=dict(imad_wide_gops=imad_gops, lop3_gops=lop3_gops,
                               per_product="144 IMAD.WIDE + 251 LOP3",
                               pipe_bound_gmul_s=pipe_bound, frac_of_pipe_bound=k_gmul / pipe_bound),
                    hbm=dict(achieved=k_gbs, peak=hbm_peak, unit="GB/s", frac=k_gbs / hbm_peak, peak_source=peak_src),
                    sumcheck_stage=dict(ms=sc_ms, share_of_step=sc_ms / sum(stage_acc.values()),
                                        algorithmic_bytes=info["sumcheck_alg_bytes"] * B,
                                        hbm_gbs=info["sumcheck_alg_bytes"] * B / (sc_ms * 1e-3) / 1e9,
                                        hbm_frac=info["sumcheck_alg_bytes"] * B / (sc_ms * 1e-3) / 1e9 / hbm_peak,
                                        gmul_s=info["sumcheck_mults"] * B / (sc_ms * 1e-3) / 1e9,
                                        frac_of_mul_peak=info["sumcheck_mults"] * B / (sc_ms * 1e-3) / 1e9 / gmul_peak,
                                        kernel_ms={k: dict(ms=v[0], launches=v[1]) for k, v in kacc.items()}),
                    stage_ms=stage_acc, top_stage=top,
                    kernels=stage_rates(info, stage_acc, B, gmul_peak, ctx.microbench(3)))

    # ---- the reference CPU prover on this box's host cores (bounded sample)
    nthreads = os.cpu_count() or 1
    cpu = cpu_reference_throughput(nthreads, 60)
    cpu1 = cpu_reference_throughput(1, 30)
    cpu["single_thread_proofs_per_s"] = cpu1["value"]
    cpu["single_thread_ms_per_proof"] = cpu1["ms_per_proof_1thread"]

    if "error" not in other.get("ecdsa_p256", {}) and not skip:
        try:
            ecpu = cpu_reference_throughput(nthreads, 6, "ecdsa")
            ecpu1 = cpu_reference_throughput(1, 4, "ecdsa")
            ecpu["single_thread_ms_per_proof"] = ecpu1["ms_per_proof_1thread"]
            other["ecdsa_p256"]["cpu_baseline"] = ecpu
            other["ecdsa_p256"]["verifier"]["cpu_1thread_ms_per_proof"] = cpu_verify_ms("ecdsa")
        except Exception as ex:
            other["ecdsa_p256"]["cpu_baseline"] = dict(error=repr(ex))

    # the SHA-256 verifier, same shape as the ECDSA one
    try:
        ver = lf.ZkVerifier(circuit)
        proofs = [bytes(h_out[i, :int(h_len[i])].numpy()) for i in range(B)]
        st, _ = ver.verify_batch(None, proofs)
        assert (st == 0).all()
        # raw-pointer timing: the C-ABI call on the host buffers the prover filled (no Python packing timed)
        import ctypes as C
        from longfellow_zk_b200 import _native
        lens64 = h_len.numpy().astype(np.uint64)
        vst, vwhy = np.zeros(B, np.int32), np.zeros(B, np.int32)
        pp = lambda a: a.ctypes.data_as(C.c_void_p)
        hout_np = h_out.numpy()
        npubb = info["npub_in"] * info["kbytes"]
        pubs = np.ascontiguousarray(h_wit.numpy()[:, :npubb]) if npubb else None

        def verify_raw():
            _native.check(_native.lib().lf_zk_verify_batch(circuit._h, B, pp(pubs) if pubs is not None else None,
                                                           pp(hout_np), pb, pp(lens64), b"test", 4, pp(vst), pp(vwhy)))
        verify_raw()
        t0 = time.perf_counter()
        for _ in range(2):
            verify_raw()
        tv = (time.perf_counter() - t0) / 2
        assert (vst == 0).all()
        sha_verifier = dict(value=B / tv, unit="proofs verified/s", ms_per_proof=1e3 * tv / B,
                            how="lf_zk_verify_batch, host buffers in / status out, batch of %d" % B,
                            cpu_1thread_ms_per_proof=cpu_verify_ms("sha"))
    except Exception as ex:
        sha_verifier = dict(error=repr(ex))

    # ---- BASELINE config 1: stand-alone FFT / RS per field with the reference's CPU time beside it
    try:
        other["config1_fft_rs"] = dict(skipped="--headline-only") if skip else measure_config1(ctx)
    except Exception as ex:
        other["config1_fft_rs"] = dict(error=repr(ex))

    # ---- the other published instance sizes (BM_ShaZK_fp2_128/2..33, BM_ECDSAZKProver/2,3)
    try:
        other["published_sizes"] = dict(skipped="--headline-only") if skip else measure_sizes(lf, ctx, stream)
    except Exception as ex:
        other["published_sizes"] = dict(error=repr(ex))

    # ---- mdoc: the reference's prover on the host cores
    try:
        from oracle import refapi
        if "error" not in other.get("mdoc", {}) and refapi.mdoc_available():
            from fixtures import load_mdoc
            mc = refapi.MdocCase(load_mdoc()["raw"])
            coins = np.random.default_rng(1).integers(0, 256, 1 << 20, dtype=np.uint8)
            mc.prove(coins)
            t0 = time.perf_counter()
            for _ in range(3):
                mc.prove(coins)
            one = (time.perf_counter() - t0) / 3

            def many():
                for _ in range(2):
                    mc.prove(coins)
            ths = [threading.Thread(target=many) for _ in range(nthreads)]
            t0 = time.perf_counter()
            for th in ths:
                th.start()
            for th in ths:
                th.join()
            wall = time.perf_counter() - t0
            other["mdoc"]["cpu_baseline"] = dict(value=2 * nthreads / wall, unit=UNIT, cores=nthreads, kind="reference",
                                                 sample="%d threads x 2 proofs (run_mdoc_prover from 'Run prover' on, "
                                                        "oracle/ref_build/ref_mdoc.cc)" % nthreads,
                                                 single_thread_ms_per_proof=1e3 * one, build=ref_build_info())
    except Exception as ex:
        other.setdefault("mdoc", {})["cpu_baseline"] = dict(error=repr(ex))

    # ---- mdoc verifier: the reference's run_mdoc_verifier, unchanged, with its two ZkVerifier objects on the
    # host (libref_mdoc_gpu.so) and resolved to ZkVerifierGpu (libref_mdoc_gpuv.so); both calls decompress and
    # parse the circuit file first, as run_mdoc_verifier does on every call (mdoc_zk.cc:549-716)
    try:
        from oracle import refapi
        if "error" not in other.get("mdoc", {}) and refapi.mdoc_gpu_available() and refapi.mdoc_gpuv_available():
            from fixtures import load_mdoc
            circuit = refapi.zstd_compress(load_mdoc()["raw"])
            A, V = refapi.mdoc_gpu_lib(), refapi.mdoc_gpuv_lib()
            code, proof = refapi.mdoc_prove_claim(V, 0, circuit)
            res = {}
            for name, L in (("reference_1thread", A), ("gpu", V)):
                rc = refapi.mdoc_verify_claim(L, 0, circuit, proof)   # warm: device copy of the circuits cached
                t0 = time.perf_counter()
                for _ in range(3):
                    rc |= refapi.mdoc_verify_claim(L, 0, circuit, proof)
                res[name + "_ms"] = 1e3 * (time.perf_counter() - t0) / 3
                res[name + "_accepted"] = (code == 0 and rc == 0)
            b1 = other["mdoc"].get("detail", {}).get("batches", [{}])[0]
            other["mdoc"]["verifier"] = dict(res, how="run_mdoc_verifier end to end (circuit file decompressed and "
                                             "parsed by the reference's code inside every call), one proof",
                                             c_abi_ms_batch1=dict(verify_hash=b1.get("ms_verify_hash"),
                                                                  verify_sig=b1.get("ms_verify_sig"),
                                                                  total=b1.get("ms_verify_total"),
                                                                  how="lf_zk_verify_committed_batch on both circuits, "
                                                                      "one transcript, circuits resident"))
    except Exception as ex:
        other.setdefault("mdoc", {})["verifier"] = dict(error=repr(ex))

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=ms_total / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="gf2^128 (u32 limbs)", data="synthetic",
                config=dict(workload=WORKLOAD, proofs_per_step_per_gpu=B, parallelism=f"independent proofs x{world}",
                            distinct_witnesses=ndistinct,
                            l2="inputs+working set of one step (>%d MB) exceed the 126 MB L2" %
                               (B * (wb + rb + pb) // (1 << 20)),
                            ninputs=info["ninputs"], nterms=info["nterms"], tableau=[info["nrow"], info["block_enc"]],
                            proof_bytes=int(lens[0])),
                clocks=clocks,
                streams=dict(n=NS, note="steps go round robin over NS contexts/streams; the interval is one "
                                        "CUDA-event pair on the first stream enclosing all launches of all",
                             one_stream=dict(value=world * B * args.steps / (ms_one_stream * 1e-3),
                                             ms_per_step=ms_one_stream / args.steps)),
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=B * (wb + rb),
                         d2h_bytes_per_step=B * (((int(lens.max()) + 15) & ~15) + 12), ms_per_step=ms_e2e_step,
                         how="lf_zk_prove_batch with pinned host buffers; NS batches in flight (one host "
                             "thread, context and stream each) so PCIe copies overlap the other batches' kernels",
                         one_batch_at_a_time=dict(value=e2e_serial, ms_per_step=ms_e2e / e2e_steps)),
                gpu_launches=launches, roofline=roofline, cpu_baseline=cpu,
                latency_ms_per_proof_batch1=lat_ms, verifier=sha_verifier,
                ms_per_proof=ms_total / args.steps / B, other_workloads=other)
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1024, help="independent proofs per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--headline-only", action="store_true",
                    help="skip other_workloads and the CPU legs (used for the ncu launch list of the timed path)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
