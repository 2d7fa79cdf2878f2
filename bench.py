#!/usr/bin/env python
"""Benchmark of the InfiniteVL hybrid-attention hot path on B200 (contract: see README / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--seq T] [--impl ours|reference]

A *step* is one pass of the hot path over one synthetic 128K-token sequence of the InfiniteVL-3B
decoder: the token mixers of all 36 layers (27 Gated DeltaNet chunk-prefill calls, H=16 K=128 V=256,
and 9 sliding-window attention calls, Hq=16 Hkv=2 D=128 W=8192), inputs resident in HBM.
`value` = tokens / second through that path.  With N > 1 ranks the sequence is sharded by contiguous
token range and every layer hands its DeltaNet state (and SWA K/V halo) to the next rank (strong scaling).

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_GDN_LAYERS = 27
N_SWA_LAYERS = 9
H, K, V = 16, 128, 256
HQ, HKV, D, WINDOW = 16, 2, 128, 8192
GDN_BYTES_PER_TOKEN = 24672          # q,k,v,g,beta read + o written, per token per layer
GDN_STATE_BYTES = 2 * H * K * V * 4  # h0 read + hT written, per sequence per layer
# dram__bytes_read.sum + dram__bytes_write.sum of ONE overlapped ivl_gdn_chunk_fwd call at T = 131072 (prep + scan
# running concurrently, 24-chunk L2 image ring), from the committed ncu range-replay capture
# profiles/r02f_range_ring24.csv (2 452 369 664 read + 1 224 527 872 written; without the ring 7.97 GB,
# profiles/r02d_range_overlapped.csv; round 1: 10.98 GB); refreshed whenever the kernels change
GDN_DRAM_TRAFFIC_NCU = 3676897536


def swa_flops(T, Tk_prefix=0):
    """4 * Hq * D * sum_t min(t + 1 + prefix, W)   (SURVEY.md 8d)."""
    total = 0
    lo = Tk_prefix
    # sum over t in [0, T) of min(t + 1 + lo, W)
    full_from = max(0, WINDOW - 1 - lo)          # first t with t + 1 + lo >= W
    if full_from >= T:
        total = T * (lo + 1) + T * (T - 1) // 2
    else:
        total = full_from * (lo + 1) + full_from * (full_from - 1) // 2 + (T - full_from) * WINDOW
    return 4 * HQ * D * total


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        with open(self.path) as f:
            for line in f:
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class HotPath:
    """Device-resident synthetic inputs of one rank's token range + the launches of one step."""

    def __init__(self, T_local, rank, world, device, seed=0):
        from inputs import gdn_inputs

        from infinitevl_b200 import _lib, ops
        self.lib = _lib.load()
        self._lib_mod = _lib
        self.ops = ops
        self.T = T_local
        self.rank, self.world = rank, world
        self.dev = device
        gen_T = min(T_local, 16384)  # generate on CPU in pieces, tile to length (values are i.i.d. anyway)
        q, k, v, g, beta, h0 = gdn_inputs(T=gen_T, H=H, seed=seed + rank)
        rep = (T_local + gen_T - 1) // gen_T
        tile = lambda x: x.repeat(1, rep, *([1] * (x.dim() - 2)))[:, :T_local].contiguous().to(device)
        self.q, self.k, self.v, self.g, self.beta = (tile(x) for x in (q, k, v, g, beta))
        self.h0 = h0.to(device)
        self.state_in = torch.empty_like(self.h0)
        self.ht = torch.empty_like(self.h0)
        self.o = torch.empty(1, T_local, H, V, dtype=torch.bfloat16, device=device)
        self.ws = ops.gdn_workspace(1, T_local, H, device)
        self.launches_per_step = 0
        self.has_swa = False
        # sequence-sharded run: every layer owns its hand-off buffers, as every layer of the model owns its cache
        # (no reuse hazards between the transfers of neighbouring layers)
        from infinitevl_b200 import dist as ivl_dist
        self.ivl_dist = ivl_dist
        self.ho = ivl_dist.OperatorHandOff(rank, world, dry=os.environ.get("IVL_BENCH_NOCOMM", "0") == "1") if world > 1 else None
        if world > 1:
            self.state_ins = [torch.empty_like(self.h0) for _ in range(N_GDN_LAYERS)]
            self.hts = [torch.empty_like(self.h0) for _ in range(N_GDN_LAYERS)]
        try:
            from infinitevl_b200 import swa
            self.swa = swa
            gen = torch.Generator().manual_seed(seed + 100 + rank)
            mk = lambda h: torch.randn(1, gen_T, h, D, generator=gen).bfloat16().repeat(1, rep, 1, 1)[:, :T_local] \
                .contiguous().to(device)
            self.sq, self.sk, self.sv = mk(HQ), mk(HKV), mk(HKV)   # [B, T, H, D] (kernel-native layout)
            if world > 1:
                # [halo of W-1 keys from the previous rank ; local keys] in ONE buffer per layer: the halo is received
                # straight into the front rows, the kernel reads the whole buffer -- no concatenation
                Hh = WINDOW - 1
                self.kbufs = [torch.empty(1, Hh + T_local, HKV, D, dtype=torch.bfloat16, device=device)
                              for _ in range(N_SWA_LAYERS)]
                self.vbufs = [torch.empty_like(b) for b in self.kbufs]
                for kb, vb in zip(self.kbufs, self.vbufs):
                    kb[:, Hh:].copy_(self.sk)
                    vb[:, Hh:].copy_(self.sv)
                    kb[:, :Hh].zero_()
                    vb[:, :Hh].zero_()
            self.so = torch.empty(1, T_local, HQ, D, dtype=torch.bfloat16, device=device)
            self.has_swa = True
        except ImportError:
            self.has_swa = False
        # neighbour hand-off through peer memory (dist.PeerLink: copy engines + stream-ordered flags, no NCCL kernels);
        # IVL_SHARD_TRANSPORT=nccl keeps the isend / irecv path
        self.link = None
        self.transport = "nccl"
        if world > 1 and os.environ.get("IVL_SHARD_TRANSPORT", "p2p") == "p2p" and os.environ.get("IVL_BENCH_NOCOMM", "0") != "1":
            try:
                Hh = WINDOW - 1
                n_in = min(Hh, rank * T_local)
                boxes = {f"S{g}": self.state_ins[g] for g in range(N_GDN_LAYERS)}
                if self.has_swa:
                    for i in range(N_SWA_LAYERS):
                        boxes[f"K{i}"] = self.kbufs[i][:, Hh - n_in:Hh] if n_in else None
                        boxes[f"V{i}"] = self.vbufs[i][:, Hh - n_in:Hh] if n_in else None
                link = ivl_dist.PeerLink(rank, world, None, device)
                link.open(boxes)
                self.link, self.transport = link, "p2p"
                self.ho.dry = True
            except Exception as e:  # noqa: BLE001  (no IPC / no peer access on this box: NCCL carries the hand-off)
                print(f"[bench] peer-memory hand-off unavailable ({type(e).__name__}: {e}); using NCCL", file=sys.stderr)

    # -- single kernels (for the roofline timing) ---------------------------------------------
    def gdn_prep(self):
        st = torch.cuda.current_stream().cuda_stream
        self._lib_mod.check(self.lib.ivl_gdn_chunk_prep(
            self.q.data_ptr(), self.k.data_ptr(), self.v.data_ptr(), self.g.data_ptr(), self.beta.data_ptr(),
            1, self.T, H, 0.0, 1, self.ws.data_ptr(), self.ws.numel(), st), "ivl_gdn_chunk_prep")

    def gdn_scan(self, h0):
        st = torch.cuda.current_stream().cuda_stream
        self._lib_mod.check(self.lib.ivl_gdn_chunk_scan(
            self.v.data_ptr(), h0.data_ptr(), 0, self.o.data_ptr(), self.ht.data_ptr(), 0, 1, self.T, H,
            self.ws.data_ptr(), self.ws.numel(), st), "ivl_gdn_chunk_scan")

    def gdn_fwd(self, h0):
        """The chunk operator as the model calls it (ivl_gdn_chunk_fwd): at this length prep and scan run
        overlapped on two streams, the scan following prep's per-chunk ready flags."""
        st = torch.cuda.current_stream().cuda_stream
        self._lib_mod.check(self.lib.ivl_gdn_chunk_fwd(
            self.q.data_ptr(), self.k.data_ptr(), self.v.data_ptr(), self.g.data_ptr(), self.beta.data_ptr(),
            h0.data_ptr(), 0, self.o.data_ptr(), self.ht.data_ptr(), 0, 1, self.T, H, K, V, 0.0, 1,
            self.ws.data_ptr(), self.ws.numel(), st), "ivl_gdn_chunk_fwd")

    def gdn_fwd_into(self, h0, ht):
        st = torch.cuda.current_stream().cuda_stream
        self._lib_mod.check(self.lib.ivl_gdn_chunk_fwd(
            self.q.data_ptr(), self.k.data_ptr(), self.v.data_ptr(), self.g.data_ptr(), self.beta.data_ptr(),
            h0.data_ptr(), 0, self.o.data_ptr(), ht.data_ptr(), 0, 1, self.T, H, K, V, 0.0, 1,
            self.ws.data_ptr(), self.ws.numel(), st), "ivl_gdn_chunk_fwd")

    def gdn_scan_into(self, h0, ht):
        st = torch.cuda.current_stream().cuda_stream
        self._lib_mod.check(self.lib.ivl_gdn_chunk_scan(
            self.v.data_ptr(), h0.data_ptr(), 0, self.o.data_ptr(), ht.data_ptr(), 0, 1, self.T, H,
            self.ws.data_ptr(), self.ws.numel(), st), "ivl_gdn_chunk_scan")

    def _post_recv(self, layer):
        """Post the receive of `layer`'s incoming hand-off (one layer ahead of its use)."""
        if self.ho is None or self.ho.first or layer >= N_GDN_LAYERS + N_SWA_LAYERS:
            return None
        if layer % 4 == 0:
            if not self.has_swa:
                return None
            i = layer // 4
            Hh = WINDOW - 1
            n_in = min(Hh, self.rank * self.T)     # tokens of the window that live on earlier ranks
            return self.ho.post_recv([self.kbufs[i][:, Hh - n_in:Hh], self.vbufs[i][:, Hh - n_in:Hh]])
        return self.ho.post_recv([self.state_ins[layer - layer // 4 - 1]])

    def step(self):
        """One pass of the hot path over this rank's token range.  Sharded (world > 1): the neighbour hand-off of
        infinitevl_b200.dist (OperatorHandOff / gdn_layer_sharded): receives posted one layer ahead, the GDN
        pre-pass runs before its state is awaited, sends never block the compute stream."""
        n = 0
        L = N_GDN_LAYERS + N_SWA_LAYERS
        nxt = self._post_recv(0)
        trace = getattr(self, "layer_trace", None)   # developer probe: one event per layer (IVL_BENCH_LAYER_TRACE=1)
        if trace is not None:
            trace.append([])
        for layer in range(L):
            if trace is not None:
                ev = torch.cuda.Event(enable_timing=os.environ.get("IVL_BENCH_LAYER_TRACE") != "3")
                ev.record()
                trace[-1].append(ev)
            cur, nxt = nxt, self._post_recv(layer + 1)
            if layer % 4 == 0:
                if self.has_swa:
                    n += self.swa_layer(layer // 4, cur)
                continue
            if self.world == 1:
                self.gdn_fwd(self.h0)
                n += 2
                continue
            g = layer - layer // 4 - 1
            if self.link is not None:
                lk, name = self.link, f"S{g}"
                lk.wait(name)
                lk.before_overwrite(name)
                self.gdn_fwd_into(self.h0 if self.rank == 0 else self.state_ins[g], self.hts[g])
                lk.release(name)
                lk.send(name, self.hts[g])
                n += 2
                continue
            overlap = os.environ.get("IVL_SHARD_GDN", "overlap") == "overlap"
            self.ivl_dist.gdn_layer_sharded(self.ho, self.gdn_prep, lambda h0, g=g: self.gdn_scan_into(h0, self.hts[g]),
                                            self.h0, self.state_ins[g], self.hts[g], cur,
                                            fwd=(lambda h0, g=g: self.gdn_fwd_into(h0, self.hts[g])) if overlap else None)
            n += 2
        if self.ho is not None:
            self.ho.drain()
        if self.link is not None:
            self.link.drain()
        self.launches_per_step = n
        return n

    def swa_layer(self, i=0, pending=None):
        if self.world > 1:
            Hh = WINDOW - 1
            kb, vb = self.kbufs[i], self.vbufs[i]
            n_in = min(Hh, self.rank * self.T)
            n_out = min(Hh, (self.rank + 1) * self.T)
            if self.link is not None:
                lk = self.link
                send = lambda: (lk.send(f"K{i}", kb[:, Hh + self.T - n_out:]), lk.send(f"V{i}", vb[:, Hh + self.T - n_out:]))
                if n_out <= self.T:
                    send()
                if n_in:
                    lk.wait(f"K{i}")
                    lk.wait(f"V{i}")
                if n_out > self.T:
                    send()
                self.swa.swa_attention_bthd(self.sq, kb[:, Hh - n_in:], vb[:, Hh - n_in:], window=WINDOW, out=self.so,
                                            key_pos0=self.rank * self.T - n_in)
                if n_in:
                    lk.release(f"K{i}")
                    lk.release(f"V{i}")
                return 1
            # halo hand-off: the last W-1 keys/values up to the end of this rank's range go to the next rank (the tail
            # of the layer's K/V buffer: contiguous, no staging copy).  Only when the local range is shorter than the
            # window does the outgoing halo contain received rows, and the send has to follow the receive.
            if n_out <= self.T:
                self.ho.post_send([kb[:, Hh + self.T - n_out:], vb[:, Hh + self.T - n_out:]])
            if pending is not None:
                pending.wait()
            if n_out > self.T:
                self.ho.post_send([kb[:, Hh + self.T - n_out:], vb[:, Hh + self.T - n_out:]])
            self.swa.swa_attention_bthd(self.sq, kb[:, Hh - n_in:], vb[:, Hh - n_in:], window=WINDOW, out=self.so,
                                        key_pos0=self.rank * self.T - n_in)
            return 1
        self.swa.swa_attention_bthd(self.sq, self.sk, self.sv, window=WINDOW, out=self.so)
        return 1


def time_events(fn, iters):
    evs = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in evs]


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T = args.seq
    assert T % (64 * world) == 0
    T_local = T // world
    hp = HotPath(T_local, rank, world, dev)
    peaks = read_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        hp.step()
    barrier()
    if os.environ.get("IVL_BENCH_LAYER_TRACE", "0") == "1":
        hp.layer_trace = []
        t0 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(4):
            hp.step()
        torch.cuda.synchronize()
        rows = [[round(t0.elapsed_time(e), 2) for e in st] for st in hp.layer_trace]
        print(f"[layer-trace rank {rank}] start of every layer (ms since the loop start), steps 0..3:\n" +
              "\n".join(" ".join(f"{x:7.2f}" for x in r) for r in rows), file=sys.stderr, flush=True)
        hp.layer_trace = None
        barrier()
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local_rank)
    if rank == 0 and os.environ.get("IVL_BENCH_NO_SAMPLER", "0") != "1":
        sampler.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if os.environ.get("IVL_BENCH_LAYER_TRACE", "0") in ("2", "3"):
        hp.layer_trace = []
    barrier()
    a.record()
    for _ in range(args.steps):
        hp.step()
    b.record()
    barrier()
    if getattr(hp, "layer_trace", None) and os.environ.get("IVL_BENCH_LAYER_TRACE") == "2":
        rows = [[round(a.elapsed_time(e), 2) for e in st] for st in hp.layer_trace]
        print(f"[layer-trace rank {rank}] timed loop, start of every layer (ms):\n" +
              "\n".join(" ".join(f"{x:7.2f}" for x in r) for r in rows) + f"\n end {a.elapsed_time(b):.2f}", file=sys.stderr, flush=True)
        hp.layer_trace = None
    ms_total = torch.tensor([a.elapsed_time(b)], device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total.item() / args.steps

    # ---- sharded runs: latency of ONE prompt (the loop above pipelines prompts: rank 0 starts step i + 1 while the
    #      last rank still finishes step i, so the wavefront fill of P - 1 layers is paid once per loop) and parity
    #      of the package's sharded prefill against the one-GPU run
    dist_info = None
    if world > 1:
        lat = []
        for _ in range(max(3, min(args.steps, 5))):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            hp.step()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            lat.append(t.item())
        lat.sort()
        dist_info = {"transport": hp.transport, "single_prompt_ms": round(lat[len(lat) // 2], 3),
                     "single_prompt_tokens_per_s": round(T / (lat[len(lat) // 2] * 1e-3), 1),
                     "pipelined_ms_per_step": round(ms_step, 3),
                     "ideal_wavefront_efficiency": round((N_GDN_LAYERS + N_SWA_LAYERS) / (N_GDN_LAYERS + N_SWA_LAYERS + world - 1), 4)}
        if not args.no_parity:
            Tp = min(T, 32768) if T % (64 * world) == 0 else 64 * world * 8
            op = hp.ivl_dist.operator_parity_check(T=Tp, transport=hp.transport)
            par = hp.ivl_dist.sharded_parity_check(T=Tp, num_layers=8)
            keys = ("out", "state", "kv", "conv", "oneshot_out", "oneshot_out_chunked_on_one_gpu", "oneshot_state")
            par_t = torch.zeros(len(keys) + 2, device=dev)
            if par:
                par_t = torch.tensor([par[k] for k in keys] + [1.0 if par["ints_equal"] else 0.0,
                                                                 1.0 if par["bit_identical"] else 0.0], device=dev)
            dist.all_reduce(par_t, op=dist.ReduceOp.MAX)
            pl = par_t.tolist()
            dist_info["parity_err"] = {
                "operators": op,
                "decoder": dict({k: pl[i] for i, k in enumerate(keys)}, ints_equal=bool(pl[-2]), bit_identical=bool(pl[-1])),
                "what": "operators: the two hot-path operators with the NCCL neighbour hand-off vs the one-shot call on the "
                        "whole sequence (T=%d), bit for bit.  decoder: dist.sharded_prefill (HybridDecoder, 8 layers, 3B "
                        "mixer dims) vs the same token ranges run one after the other on one GPU (out/state/kv/conv: RMS "
                        "error ratio, must be 0) and vs one call over the whole sequence (oneshot_*: gate 1e-3, BASELINE.md 3c)" % Tp}

    # ---- per-kernel timing for the roofline (device events on the launching stream) --------------
    roof = None
    kernels = {}
    if rank == 0:
        prep = time_events(hp.gdn_prep, 10)
        scan = time_events(lambda: hp.gdn_scan(hp.h0), 10)
        fwd = time_events(lambda: hp.gdn_fwd(hp.h0), 10)
        t_prep, t_scan, t_fwd = sum(prep) / len(prep), sum(scan) / len(scan), sum(fwd) / len(fwd)
        shard_overlap = os.environ.get("IVL_SHARD_GDN", "overlap") == "overlap"
        t_layer = t_fwd if (world == 1 or shard_overlap) else t_prep + t_scan   # what step() launches per GDN layer
        alg_bytes = GDN_BYTES_PER_TOKEN * T_local + GDN_STATE_BYTES
        achieved = alg_bytes / (t_layer * 1e-3) / 1e9
        roof = {"bound": "hbm",
                "kernel": "gdn_chunk = gdn_prep_kernel + gdn_scan_t3_kernel (one GDN layer; "
                          + ("overlapped on two streams, timed as one operator call" if (world == 1 or shard_overlap) else "back to back") + ")",
                "achieved": round(achieved, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": round(achieved / peaks["hbm_gbs"], 4), "traffic": GDN_DRAM_TRAFFIC_NCU if (T_local == 131072 and world == 1) else None,
                "traffic_note": "ncu --replay-mode range over one overlapped operator call (kernel replay would serialise "
                                "the two kernels): prep reads q, k, g, beta and hands 64 KiB of operand images per chunk and head to "
                                "the scan through a 24-chunk ring that stays in L2; the scan also reads v and writes o; see "
                                "profiles/r02_summary.md",
                "peak_source": peaks["source"], "algorithmic_bytes_per_launch": alg_bytes}
        kernels = {"gdn_layer_ms": round(t_layer, 4), "gdn_prep_alone_ms": round(t_prep, 4),
                   "gdn_scan_alone_ms": round(t_scan, 4)}
        if hp.has_swa:
            swa_t = time_events(hp.swa_layer, 5) if world == 1 else None
            if swa_t:
                t_swa = sum(swa_t) / len(swa_t)
                kernels["swa_fwd_ms"] = round(t_swa, 4)
                kernels["swa_tflops"] = round(swa_flops(T_local) / (t_swa * 1e-3) / 1e12, 1)
                kernels["swa_frac_of_bf16_sustained"] = round(kernels["swa_tflops"] / peaks["bf16_tflops_sustained"], 4)

    # ---- end-to-end through the public operator API with HOST buffers ---------------------------
    decode = None
    e2e = run_e2e(hp, args, world, dev)
    gpu_ref = config2 = config3 = None
    if world == 1 and hp.has_swa:
        decode = run_decode(dev, peaks)
        if not args.no_gpu_reference:
            gpu_ref = gpu_reference(hp, T, ms_step)
        if T != 32768 and not args.no_config2:
            config2 = run_config2(dev, peaks)
        if not args.no_config3:
            config3 = run_config3(dev, peaks, n_frames=args.stream_frames)
    backward = None
    if world == 1 and not args.no_backward:
        backward = run_backward(dev)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(sample_T=args.cpu_sample, has_swa=hp.has_swa)

    if rank == 0:
        workload = (f"prefill hot path, T={T} tokens, InfiniteVL-3B mixers: {N_GDN_LAYERS}x GDN chunk (H16 K128 V256)"
                    + (f" + {N_SWA_LAYERS}x SWA (Hq16 Hkv2 D128 W8192)" if hp.has_swa else " (SWA kernel not built: GDN layers only)"))
        line = {
            "metric": "prefill tokens/sec @128K seq InfiniteVL-3B hybrid-attention hot path",
            "value": round(T / (ms_step * 1e-3), 1), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload, "seq_len": T, "batch": 1,
                       "parallelism": (f"sequence-chunk x{world}, neighbour hand-off over " + ("peer memory (CUDA IPC mapping, ivl_peer_put stores over NVLink + stream-ordered flags)" if hp.transport == "p2p" else "NCCL send/recv")) if world > 1 else "single GPU",
                       "l2": "inputs (>3 GB per layer) exceed the 126 MB L2; no flush needed"},
            "roofline": roof, "kernels": kernels, "cpu_baseline": cpu, "e2e": e2e, "decode": decode,
            "gpu_reference": gpu_ref, "config2_32k": config2, "config3_stream": config3, "backward": backward,
            "dist": dist_info,
            "gpu_launches": hp.launches_per_step * args.steps, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def gpu_reference(hp, T, ours_ms_step, warmup=3, iters=3):
    """The reference's own GPU path on the same box, same process, same inputs: pip flash-linear-attention's Triton
    chunk_gated_delta_rule (requirements.txt:19-20; called at modeling_infinitevl.py:1298-1308) and flash-attn's
    flash_attn_func with the sliding window (requirements.txt:18; modeling_infinitevl.py:1092-1108), as one step of
    27 + 9 calls.  Library code, timed as a denominator only -- nothing of it is on our path."""
    res = {"what": "fla Triton chunk_gated_delta_rule x27 + flash_attn_func(window) x9 on the same inputs, same process"}
    try:
        import fla
        import flash_attn
        from fla.ops.gated_delta_rule import chunk_gated_delta_rule as fla_chunk
        from flash_attn import flash_attn_func
        res["versions"] = {"fla": getattr(fla, "__version__", "?"), "flash_attn": getattr(flash_attn, "__version__", "?")}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"import failed: {e!r}"}

    def gdn():
        return fla_chunk(hp.q, hp.k, hp.v, hp.g, hp.beta, initial_state=hp.h0, output_final_state=True,
                         use_qk_l2norm_in_kernel=True)

    def swa():
        return flash_attn_func(hp.sq, hp.sk, hp.sv, causal=True, window_size=(WINDOW - 1, WINDOW - 1))

    def step():
        for layer in range(N_GDN_LAYERS + N_SWA_LAYERS):
            if layer % 4 == 0:
                swa()
            else:
                gdn()

    try:
        for _ in range(warmup):
            gdn()
            swa()
        torch.cuda.synchronize()
        t_gdn = sorted(time_events(gdn, 5))[2]
        t_swa = sorted(time_events(swa, 5))[2]
        step()
        torch.cuda.synchronize()
        ts = sorted(time_events(step, iters))
        ms = ts[len(ts) // 2]
    except Exception as e:  # noqa: BLE001   (OOM at very long T is a result, not a failure of the bench)
        torch.cuda.empty_cache()
        return {**res, "unavailable": f"reference GPU path failed at T={T}: {type(e).__name__}: {str(e)[:200]}"}
    res.update({"ms_per_step": round(ms, 3), "tokens_per_s": round(T / (ms * 1e-3), 1), "gdn_layer_ms": round(t_gdn, 4),
                "swa_layer_ms": round(t_swa, 4), "ours_ms_per_step": round(ours_ms_step, 3),
                "hot_path_ratio": round(ms / ours_ms_step, 3)})
    return res


def run_backward(dev, T=4096):
    """forward + backward of the chunk operator through ops.ChunkGatedDeltaRuleFunction (CUDA forward,
    ivl_gdn_bwd: exact fp32 gradient with checkpoint recomputation).  Reported, not optimised: the backward is the
    parity anchor of the training drop-in."""
    try:
        from inputs import gdn_inputs
        from infinitevl_b200 import ops
        q, k, v, g, beta, h0 = gdn_inputs(T=T, H=H, seed=11)
        leaves = [x.to(dev).requires_grad_(True) for x in (q, k, v, g, beta)]

        def fwd_bwd():
            for x in leaves:
                x.grad = None
            o, _ = ops.chunk_gated_delta_rule(*leaves, use_qk_l2norm_in_kernel=True)
            o.float().square().mean().backward()

        def fwd():
            with torch.no_grad():
                ops.chunk_gated_delta_rule(*leaves, use_qk_l2norm_in_kernel=True)
        tf = sorted(time_events(fwd, 3))[1]
        tb = sorted(time_events(fwd_bwd, 3))[1]
        return {"seq_len": T, "forward_ms": round(tf, 3), "forward_backward_ms": round(tb, 3),
                "what": "one GDN layer (H16 K128 V256), chunk forward + exact fp32 recurrence backward (ivl_gdn_bwd)"}
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {str(e)[:200]}"}


def run_config2(dev, peaks, T=32768):
    """BASELINE.json config 2: the same hot path at 32K tokens on one GPU (throughput + roofline of the GDN operator)."""
    hp = HotPath(T, 0, 1, dev, seed=7)
    for _ in range(3):
        hp.step()
    torch.cuda.synchronize()
    ts = sorted(time_events(hp.step, 5))
    ms = ts[len(ts) // 2]
    fwd = sorted(time_events(lambda: hp.gdn_fwd(hp.h0), 10))
    t_layer = fwd[len(fwd) // 2]
    alg = GDN_BYTES_PER_TOKEN * T + GDN_STATE_BYTES
    out = {"seq_len": T, "ms_per_step": round(ms, 3), "tokens_per_s": round(T / (ms * 1e-3), 1),
           "gdn_layer_ms": round(t_layer, 4), "gdn_achieved_gbs": round(alg / (t_layer * 1e-3) / 1e9, 1),
           "gdn_hbm_frac": round(alg / (t_layer * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)}
    if hp.has_swa:
        sw = sorted(time_events(hp.swa_layer, 5))
        t_swa = sw[len(sw) // 2]
        out.update({"swa_fwd_ms": round(t_swa, 4), "swa_tflops": round(swa_flops(T) / (t_swa * 1e-3) / 1e12, 1)})
    del hp
    torch.cuda.empty_cache()
    return out


def run_sweep(args):
    """BASELINE.json config 5: prefill throughput of the hot path for T = 4K ... 1M at this world size, next to the
    reference's GPU path (one GPU only: the reference does not shard a sequence).  Writes one JSON file."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rows = []
    for T in [int(x) for x in args.sweep.split(",")]:
        if T % (64 * world):
            continue
        row = {"seq_len": T, "n_gpus": world}
        try:
            hp = HotPath(T // world, rank, world, dev)
            for _ in range(2):
                hp.step()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                hp.step()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / 3], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            row.update({"ms_per_step": round(t.item(), 3), "tokens_per_s": round(T / (t.item() * 1e-3), 1)})
            if world == 1 and not args.no_gpu_reference:
                row["gpu_reference"] = gpu_reference(hp, T, t.item(), warmup=2, iters=2)
            del hp
        except Exception as e:  # noqa: BLE001
            row["error"] = f"{type(e).__name__}: {str(e)[:200]}"
        from infinitevl_b200 import ops
        ops.release_workspaces()
        torch.cuda.empty_cache()
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"sweep_n{world}.json"), "w") as f:
            json.dump({"workload": "prefill hot path (27 GDN + 9 SWA mixer calls), synthetic", "rows": rows}, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


def run_config3(dev, peaks, n_frames=2048, frame=256, decode_steps=256):
    """BASELINE.json config 3: stream n_frames x 256-token frames (= 524 288 tokens) through the
    36-layer decoder stack with the inference cache -- ONE captured CUDA graph replayed per frame, as
    inference_examples/demo_streaming_inference.py:453-489 does -- assert that memory stays flat, then time
    single-token decode steps under a graph from the state the stream left: (i) token mixers only (the hot path with
    its projections), (ii) the whole decoder (norms + MLPs too).  Random-init weights of the 3B shapes."""
    from infinitevl_b200 import modeling, ops
    res = {"frames": n_frames, "frame_tokens": frame, "context_tokens": n_frames * frame}
    for name, mixers_only in (("mixers_only", True), ("whole_decoder", False)):
        torch.manual_seed(0)
        cfg = modeling.HybridTextConfig()
        dec = modeling.HybridDecoder(cfg, mixers_only=mixers_only)
        for p_ in dec.parameters():
            if p_.dim() >= 2:
                torch.nn.init.normal_(p_, std=0.02)
        dec = dec.to(dev, torch.bfloat16).eval()
        cache = dec.allocate_inference_cache(1)
        gen = torch.Generator().manual_seed(3)
        x = torch.randn(1, frame, cfg.hidden_size, generator=gen).bfloat16().to(dev)
        sx = x.clone()
        spos = torch.zeros(3, 1, frame, dtype=torch.long, device=dev)
        scp = torch.zeros(frame, dtype=torch.long, device=dev)
        ar = torch.arange(frame, device=dev)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ops.gdn_workspace(1, frame, cfg.num_linear_heads, dev)
            for i in range(2):      # eager warm-up frames (also start every cache layer)
                scp.copy_(ar + i * frame); spos.copy_(scp[None, None].expand(3, 1, -1))
                dec(sx, position_ids=spos, past_key_values=cache, cache_position=scp)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                out = dec(sx, position_ids=spos, past_key_values=cache, cache_position=scp)
        torch.cuda.current_stream().wait_stream(side)
        mem10 = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(2, n_frames):
            if i == 12:
                torch.cuda.synchronize()
                mem10 = torch.cuda.memory_allocated()
                e0.record()
            scp.copy_(ar + i * frame); spos.copy_(scp[None, None].expand(3, 1, -1))
            sx.copy_(x)
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        frame_ms = e0.elapsed_time(e1) / (n_frames - 12)
        flat = torch.cuda.memory_allocated() == mem10
        assert flat, "streaming grew device memory"
        assert torch.isfinite(out).all()
        # ---- decode steps from the streamed state
        tok = torch.randn(1, 1, cfg.hidden_size, generator=gen).bfloat16().to(dev)
        dpos = torch.zeros(3, 1, 1, dtype=torch.long, device=dev)
        dcp = torch.zeros(1, dtype=torch.long, device=dev)
        T0 = n_frames * frame
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(2):
                dcp.fill_(T0 + i); dpos.fill_(T0 + i)
                dec(tok, position_ids=dpos, past_key_values=cache, cache_position=dcp)
            gd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gd, stream=side):
                dout = dec(tok, position_ids=dpos, past_key_values=cache, cache_position=dcp)
        torch.cuda.current_stream().wait_stream(side)
        ts = []
        for i in range(decode_steps):
            dcp.fill_(T0 + 2 + i); dpos.fill_(T0 + 2 + i)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gd.replay(); b.record()
            ts.append((a, b))
        torch.cuda.synchronize()
        tt = sorted(a.elapsed_time(b) for a, b in ts)
        assert torch.isfinite(dout).all()
        res[name] = {"frame_ms": round(frame_ms, 4), "frame_tokens_per_s": round(frame / (frame_ms * 1e-3), 1),
                     "decode_step_ms_median": round(tt[len(tt) // 2], 4), "memory_flat": bool(flat),
                     "memory_allocated_mb": round(torch.cuda.memory_allocated() / 2**20, 1)}
        del dec, cache, g, gd, out, dout
        torch.cuda.empty_cache()
    res["what"] = ("HybridDecoder (36 layers, 3B shapes, random init) with the ring-buffer / static inference cache; one "
                   "CUDA graph per frame shape, replayed; decode = 256 single-token graph replays after the stream")
    return res


def gdn_mixer_core_decode(dev, steps=50):
    """Everything between the input projections and o_proj of the 27 GDN mixers for one decode token (conv steps,
    gates, recurrence, gated norm, cache updates): the kernel-by-kernel chain against the one-launch fused step
    (ivl_gdn_decode_step), each as one CUDA graph of 27 layers."""
    from infinitevl_b200 import _lib, modeling, ops
    lib = _lib.load()
    gen = torch.Generator().manual_seed(8)
    rnd = lambda *shape: (torch.randn(*shape, generator=gen) * 0.5).bfloat16().to(dev)
    xq, xk, xv, a, b, gate = rnd(1, 1, H * K), rnd(1, 1, H * K), rnd(1, 1, H * V), rnd(1, 1, H), rnd(1, 1, H), rnd(1, 1, H * V)
    wq, wk, wv, nw = rnd(H * K, 1, 4), rnd(H * K, 1, 4), rnd(H * V, 1, 4), torch.ones(V, dtype=torch.bfloat16, device=dev)
    A_log = torch.zeros(H, device=dev)
    dt_bias = torch.zeros(H, device=dev)
    layers = [dict(cq=rnd(1, H * K, 4), ck=rnd(1, H * K, 4), cv=rnd(1, H * V, 4), st=rnd(1, H, K, V),
                   cq2=rnd(1, H * K, 4), ck2=rnd(1, H * K, 4), cv2=rnd(1, H * V, 4)) for _ in range(N_GDN_LAYERS)]
    out = torch.empty(1, 1, H * V, dtype=torch.bfloat16, device=dev)

    def chain():
        for L in layers:
            q, _ = modeling.short_conv_silu(xq, wq, L["cq"], True, cache_out=L["cq2"])
            k, _ = modeling.short_conv_silu(xk, wk, L["ck"], True, cache_out=L["ck2"])
            v, _ = modeling.short_conv_silu(xv, wv, L["cv"], True, cache_out=L["cv2"])
            L["cq"].copy_(L["cq2"]); L["ck"].copy_(L["ck2"]); L["cv"].copy_(L["cv2"])   # the cache's "set"
            g, beta = modeling.gdn_gates(a, b, A_log, dt_bias)
            o, _ = ops.fused_recurrent_gated_delta_rule(q.view(1, 1, H, K), k.view(1, 1, H, K), v.view(1, 1, H, V), g, beta,
                                                        initial_state=L["st"], output_final_state=True,
                                                        use_qk_l2norm_in_kernel=True, state_out=L["st"])
            modeling.rmsnorm_gated(o, gate.view(1, 1, H, V), nw)

    def fused():
        st = torch.cuda.current_stream().cuda_stream
        for L in layers:
            _lib.check(lib.ivl_gdn_decode_step(
                xq.data_ptr(), xk.data_ptr(), xv.data_ptr(), a.data_ptr(), b.data_ptr(), gate.data_ptr(), wq.data_ptr(),
                wk.data_ptr(), wv.data_ptr(), A_log.data_ptr(), dt_bias.data_ptr(), nw.data_ptr(), L["cq"].data_ptr(),
                L["ck"].data_ptr(), L["cv"].data_ptr(), L["st"].data_ptr(), 1, out.data_ptr(), 1, H, K, V, 0.0, 1e-5, st),
                "ivl_gdn_decode_step")

    res = {}
    for name, fn in (("kernel_chain_ms", chain), ("fused_ms", fused)):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        for _ in range(5):
            graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        res[name] = round(e0.elapsed_time(e1) / steps, 4)
    res["what"] = ("27 GDN mixers, one decode token, projections excluded: 3 conv steps + cache copies + gates + "
                   "recurrence + gated norm per layer (10 launches) vs ivl_gdn_decode_step (1 launch)")
    return res


def run_decode(dev, peaks, context=524288, steps=50):
    """Decode step of the hot path (BASELINE.json config 3): after a 512K-token context the state is constant
    size -- 27 DeltaNet states (bf16, as the reference caches them) and 9 full 8191-token K/V windows -- so one
    step is 27 token-recurrence launches + 9 split-KV attention launches, captured in one CUDA graph."""
    from inputs import gdn_inputs

    from infinitevl_b200 import ops, swa
    q, k, v, g, beta, h0 = gdn_inputs(T=1, H=H, seed=5, device=dev)
    states = [h0.to(torch.bfloat16).clone() for _ in range(N_GDN_LAYERS)]
    gen = torch.Generator().manual_seed(6)
    from infinitevl_b200.cache import StaticSlidingWindowLayerPrealloc
    from infinitevl_b200.modeling import HybridTextConfig
    cfg = HybridTextConfig()
    rings = [StaticSlidingWindowLayerPrealloc(config=cfg, batch_size=1, device=dev, dtype=torch.bfloat16)
             for _ in range(N_SWA_LAYERS)]
    fill = torch.randn(1, WINDOW + 300, HKV, D, generator=gen).bfloat16().to(dev)
    for r in rings:      # full windows, ring wrapped once: the steady state of a long stream
        for a in range(0, fill.shape[1], 256):
            r._append(fill[:, a:a + 256], fill[:, a:a + 256])
    sq = torch.randn(1, 1, HQ, D, generator=gen).bfloat16().to(dev)
    sk = torch.randn(1, 1, HKV, D, generator=gen).bfloat16().to(dev)

    def step():
        for st in states:
            ops.fused_recurrent_gated_delta_rule(q, k, v, g, beta, initial_state=st, output_final_state=True,
                                                 use_qk_l2norm_in_kernel=True, state_out=st)
        for r in rings:      # append the token's K/V + attention over the window + combine: one launch per layer
            r.attend(sq, sk, sk, D ** -0.5, WINDOW)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            step()
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    alg = N_GDN_LAYERS * 2 * H * K * V * 2 + N_SWA_LAYERS * 2 * HKV * WINDOW * D * 2
    core = gdn_mixer_core_decode(dev)
    return {"step_ms": round(ms, 4), "gdn_mixer_core": core, "context_tokens": context, "launches_per_step": N_GDN_LAYERS + N_SWA_LAYERS,
            "algorithmic_bytes": alg, "achieved_gbs": round(alg / (ms * 1e-3) / 1e9, 1),
            "hbm_frac": round(alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
            "what": "mixer kernels only: 27x GDN token recurrence (bf16 state in place) + 9x ring-cache SWA decode step "
                    "(K/V append + split-KV attention over the 8192-token window + combine in one launch), one CUDA "
                    "graph; state size is independent of the context length.  config3_stream has the step through the "
                    "decoder modules (projections, cache classes) after a real 512K-token stream"}


def run_e2e(hp, args, world=1, dev=None):
    """Same step through the C ABI, but every step's inputs start in pinned HOST memory and its result is read
    back to the host, all inside the timed region (every rank feeds its own token range; max over ranks).
    The feeder is double-buffered: the host->device copy of step i+1 and the device->host read of step i-1 run
    on a copy stream while step i computes, which is how a caller with host-resident activations would drive
    the operators; every byte is still copied every step."""
    import torch.distributed as dist
    T = hp.T * world
    names = ["q", "k", "v", "g", "beta"] + (["sq", "sk", "sv"] if hp.has_swa else [])
    host = {n: getattr(hp, n).cpu().pin_memory() for n in names}
    outs = ["o"] + (["so"] if hp.has_swa else [])      # every result of the step leaves the device
    out_host = {n: torch.empty(getattr(hp, n).shape, dtype=getattr(hp, n).dtype).pin_memory() for n in outs}
    h2d = sum(x.numel() * x.element_size() for x in host.values())
    d2h = sum(x.numel() * x.element_size() for x in out_host.values())
    sets = [{n: getattr(hp, n) for n in names + outs},
            {n: torch.empty_like(getattr(hp, n)) for n in names + outs}]
    compute = torch.cuda.current_stream()
    copy = torch.cuda.Stream()       # host -> device
    copy_out = torch.cuda.Stream()   # device -> host: its own stream, so the two DMA directions run concurrently
    fed = [torch.cuda.Event() for _ in range(2)]      # inputs of the set are on the device
    free = [torch.cuda.Event() for _ in range(2)]     # the set's step has run: inputs may be overwritten
    read = [torch.cuda.Event() for _ in range(2)]     # the set's output has been read back

    def feed(i):
        st = sets[i % 2]
        with torch.cuda.stream(copy):
            if i >= 2:
                copy.wait_event(free[i % 2])
            for n in names:
                st[n].copy_(host[n], non_blocking=True)
            fed[i % 2].record(copy)

    def run(i):
        st = sets[i % 2]
        compute.wait_event(fed[i % 2])
        if i >= 2:
            compute.wait_event(read[i % 2])           # the previous output in this set has left the device
        for n in names + outs:
            setattr(hp, n, st[n])
        if world > 1 and hp.has_swa:     # sharded: the layers read K/V from their hand-off buffers
            Hh = WINDOW - 1
            for kb, vb in zip(hp.kbufs, hp.vbufs):
                kb[:, Hh:].copy_(st["sk"])
                vb[:, Hh:].copy_(st["sv"])
        hp.step()
        free[i % 2].record(compute)
        with torch.cuda.stream(copy_out):
            copy_out.wait_event(free[i % 2])
            for n in outs:
                out_host[n].copy_(st[n], non_blocking=True)
            read[i % 2].record(copy_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def loop(steps):
        feed(0)
        for i in range(steps):
            if i + 1 < steps:
                feed(i + 1)
            run(i)

    loop(2)
    barrier()
    steps = max(4, min(2 * args.steps, 10))   # the feeder is a two-deep pipeline: enough steps to amortise its fill
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    loop(steps)
    copy.synchronize()
    copy_out.synchronize()
    b.record()
    barrier()
    for n in names + outs:
        setattr(hp, n, sets[0][n])
    ms_t = torch.tensor([a.elapsed_time(b) / steps], device=hp.dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms = ms_t.item()
    return {"value": round(T / (ms * 1e-3), 1), "unit": "tokens/s", "h2d_bytes_per_step": h2d * world,
            "d2h_bytes_per_step": d2h * world, "ms_per_step": round(ms, 3), "steps": steps,
            "api": "infinitevl_b200 C ABI (ivl_gdn_chunk_fwd" + (", ivl_swa_fwd" if hp.has_swa else "")
                   + ") fed from pinned host buffers, double-buffered, one copy stream per direction"}


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle (a port: the reference has no CPU path for this operator)
# ------------------------------------------------------------------------------------------------
def cpu_hot_path_seconds(sample_T, has_swa=True):
    from inputs import gdn_inputs
    from oracle import gdn_chunk_ref, swa_attention_ref
    torch.set_num_threads(os.cpu_count())
    q, k, v, g, beta, h0 = gdn_inputs(T=sample_T, H=H, seed=0)
    t0 = time.perf_counter()
    gdn_chunk_ref(q, k, v, g, beta, initial_state=h0)
    t_gdn = time.perf_counter() - t0
    t_swa = 0.0
    if has_swa:
        # the sample's queries sit deep inside a long sequence: every one of them sees a full window (W - 1 cached keys
        # in front of the sample), which is the per-token cost of the 128K-token workload -- a short causal prompt
        # would see a quarter of a window on average and overstate the CPU's tokens/s
        gen = torch.Generator().manual_seed(1)
        sq = torch.randn(1, HQ, sample_T, D, generator=gen)
        sk = torch.randn(1, HKV, WINDOW - 1 + sample_T, D, generator=gen)
        sv = torch.randn(1, HKV, WINDOW - 1 + sample_T, D, generator=gen)
        t0 = time.perf_counter()
        swa_attention_ref(sq, sk, sv, window=WINDOW)
        t_swa = time.perf_counter() - t0
    return t_gdn, t_swa


def cpu_baseline(sample_T=2048, has_swa=True):
    t_gdn, t_swa = cpu_hot_path_seconds(sample_T, has_swa)
    total = N_GDN_LAYERS * t_gdn + (N_SWA_LAYERS * t_swa if has_swa else 0.0)
    return {"value": round(sample_T / total, 2), "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"oracle (fp32 torch) hot path at T={sample_T}: one GDN layer {t_gdn:.3f}s x{N_GDN_LAYERS}"
                      + (f" + one SWA layer ({sample_T} queries, each over a full {WINDOW}-key window) {t_swa:.3f}s x{N_SWA_LAYERS}" if has_swa else "")
                      + "; the reference has no CPU implementation of these operators (Triton / flash-attn only)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T = args.cpu_sample
    has_swa = os.path.exists(os.path.join(ROOT, "infinitevl_b200", "swa.py"))
    for _ in range(min(args.warmup, 1)):
        cpu_hot_path_seconds(T, has_swa)
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    tot = 0.0
    for _ in range(steps):
        t_gdn, t_swa = cpu_hot_path_seconds(T, has_swa)
        tot += N_GDN_LAYERS * t_gdn + (N_SWA_LAYERS * t_swa if has_swa else 0.0)
    sec = tot / steps
    val = round(T / sec, 2)
    cpu = {"value": val, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port",
           "sample": f"oracle hot path on {T} tokens deep inside a long sequence (GDN chunk scan; SWA queries each over a full "
                     f"{WINDOW}-key window); one layer of each kind timed, scaled by layer counts; "
                     f"wall {time.perf_counter() - t0:.1f}s"}
    line = {"impl": "reference", "metric": "prefill tokens/sec @128K seq InfiniteVL-3B hybrid-attention hot path",
            "value": val, "unit": "tokens/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": round(sec * 1e3, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"CPU oracle of the same hot path on a bounded sample (T={T})", "seq_len": T,
                       "batch": 1},
            "cpu_baseline": cpu,
            "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--seq", type=int, default=131072)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=2048)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-config2", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-config3", action="store_true")
    ap.add_argument("--no-backward", action="store_true")
    ap.add_argument("--stream-frames", type=int, default=2048)
    ap.add_argument("--sweep", default=None, nargs="?", const="4096,8192,16384,32768,65536,131072,262144,524288,1048576",
                    help="comma-separated sequence lengths: run the config-5 sweep instead of the bench line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.sweep:
        run_sweep(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
