"""NativeEngine: the whole training step on hand-written sm_100a kernels.

Where the reference's hot loop (distributedVggf.py:158-175) is ``model(inputs)`` ->
``cross_entropy`` -> ``zero_grad`` -> ``backward`` (DDP hooks all-reduce buckets) -> ``Adam.step``
-> two ``.item()`` syncs, all delegated to cuDNN / cuBLAS / c10d / ATen, this engine owns every
stage:

  input     fused augment kernel: uint8 HWC -> normalised bf16 im2col rows of the first conv
  forward   conv0 = tcgen05 GEMM on the im2col rows; conv1..12 = tcgen05 implicit GEMM (4-D TMA
            boxes, TMEM accumulators, bias+ReLU epilogue); max-pool; FC stack = swap-AB split-K
            tcgen05 GEMMs + bias/ReLU/Philox-dropout epilogue; fused cross-entropy (+ metrics)
  backward  FC wgrad/dgrad GEMMs (MN-major operands, no transposed copies), conv dgrad with the
            ReLU mask fused in the epilogue, fused ReLU+pool backward, conv wgrad split-K with
            fp32 red.add straight into the flat gradient arena
  reduce    per bucket, as soon as its last gradient is enqueued: fused pack(fp32->bf16, 1/ws) +
            one-shot / two-shot / NVLS reduction over peer memory on a high-priority side stream
  update    fused Adam/SGD on the bucket right behind its reduction (reads the reduced bf16 wire,
            writes fp32 master + bf16 shadow, zeroes the gradient range)
Nothing synchronises with the host inside a step; loss / accuracy accumulate in a device meter.
"""
from __future__ import annotations

import os

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from .. import ops
from ..config import DATA, TRAIN
from ..data.loader import FusedBatch
from ..models import layout as L
from ..models.vggf import VGGSpec, build_oracle
from ..parallel.buckets import BucketPlan, engine_bucket_plan
from ..parallel.process_group import distributed_is_initialized
from ..utils.metrics import DeviceMeter

BF16 = torch.bfloat16
F32 = torch.float32


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class PhaseTimer:
    """CUDA-event timers around the phases of a step (``--profile events``)."""

    def __init__(self, enabled: bool, nvtx: bool = False) -> None:
        self.enabled = enabled
        self.nvtx = nvtx
        self.pending: List[Tuple[str, torch.cuda.Event, torch.cuda.Event]] = []
        self.totals: Dict[str, float] = {}
        self.counts: Dict[str, int] = {}
        self._open: Dict[str, torch.cuda.Event] = {}

    def start(self, name: str) -> None:
        if self.nvtx:
            torch.cuda.nvtx.range_push(name)
        if self.enabled:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._open[name] = e

    def stop(self, name: str) -> None:
        if self.nvtx:
            torch.cuda.nvtx.range_pop()
        if self.enabled and name in self._open:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.pending.append((name, self._open.pop(name), e))

    def collect(self) -> Dict[str, float]:
        for name, a, b in self.pending:
            b.synchronize()
            self.totals[name] = self.totals.get(name, 0.0) + a.elapsed_time(b)
            self.counts[name] = self.counts.get(name, 0) + 1
        self.pending.clear()
        return {k: self.totals[k] / max(self.counts[k], 1) for k in self.totals}


class NativeEngine:
    def __init__(self, spec: VGGSpec, device: torch.device, batch: int, lr: float = TRAIN.learning_rate,
                 optimizer: str = "adam", momentum: float = TRAIN.momentum, weight_decay: float = 0.0,
                 compute_dtype: str = "bf16", allreduce: str = "auto", wire_dtype: str = "bf16",
                 bucket_mb: float = 32.0, seed: int = 0, pretrained_state: Optional[dict] = None,
                 profile: Optional[str] = None, input_hw: int = DATA.crop, comm_ctas: int = 48,
                 init_state: Optional[Dict[str, torch.Tensor]] = None, unpack_fp32: bool = False,
                 distributed: bool = True, zero1="auto", conv_bucket_mb: float = 9.5,
                 tail_bucket_kb: float = 2400.0, same_dropout_all_ranks: bool = False) -> None:
        ops.require()
        if compute_dtype != "bf16":
            raise NotImplementedError("the native engine computes in bf16 with fp32 master weights; "
                                      "use --engine oracle for fp32 numerics")
        for c in spec.convs[1:]:
            if c.cin % 64 or c.cout % 64:
                raise ValueError("native conv kernels need channel counts that are multiples of 64")
        if spec.convs[0].cout % 64 or spec.convs[0].cin != 3:
            raise ValueError("first conv must be 3 -> multiple of 64 channels")
        self.spec, self.device, self.B = spec, torch.device(device), int(batch)
        self.lr, self.opt_name, self.momentum, self.weight_decay = lr, optimizer, momentum, weight_decay
        self.beta1, self.beta2, self.eps = TRAIN.adam_beta1, TRAIN.adam_beta2, TRAIN.adam_eps
        self.seed, self.HW = int(seed), int(input_hw)
        self.same_dropout_all_ranks = bool(same_dropout_all_ranks)
        self.step_count = 0
        self.train_dropout = True        # tests switch dropout off to compare gradients exactly
        self.apply_updates = True        # False: leave raw gradients in g32 (gradient inspection)
        self.meter: Optional[DeviceMeter] = None
        self.class_weights: Optional[torch.Tensor] = None
        self.timer = PhaseTimer(profile == "events", nvtx=(profile == "nvtx"))
        self._comm_events: Optional[list] = None     # [(start, end, wire bytes)] while comm timing is on
        self._tl: Optional[list] = None              # step timeline marks (timeline())
        self.nvtx = profile == "nvtx"
        self.world = dist.get_world_size() if (distributed and distributed_is_initialized()) else 1
        self.rank = dist.get_rank() if (distributed and distributed_is_initialized()) else 0
        self.ar_algo, self.comm_ctas, self.unpack_fp32 = allreduce, comm_ctas, unpack_fp32
        self.use_nccl = allreduce == "nccl"
        dev = self.device

        # ---- flat arenas (gradient-ready order) --------------------------------------------------
        # FC weights (ready first, 88 % of the bytes): ``bucket_mb`` messages.  Convolution gradients
        # (ready one layer at a time over the rest of backward): one bucket per big layer, and a small
        # final bucket -- its reduction and update are the only exposed part of the exchange.
        self.plan: BucketPlan = engine_bucket_plan(L.ready_order(spec), bucket_mb, conv_bucket_mb, tail_bucket_kb)
        n = self.plan.total
        self.p32 = torch.zeros(n, dtype=F32, device=dev)
        self.g32 = torch.zeros(n, dtype=F32, device=dev)
        self.m32 = torch.zeros(n, dtype=F32, device=dev)
        self.v32 = torch.zeros(n, dtype=F32, device=dev) if optimizer == "adam" else None
        self.w16 = torch.zeros(n, dtype=BF16, device=dev)

        if init_state is None:
            init_state = build_oracle(spec, seed=seed, pretrained_state=pretrained_state).state_dict()
        self.import_state(init_state, sync=False)

        # ---- cross-GPU plumbing ------------------------------------------------------------------
        self.arena = None
        self.cross_group = None          # inter-node stage of the hierarchical reduction
        self.comm_mode = "single" if self.world == 1 else ("nccl" if self.use_nccl else "flat")
        # high priority by default: a ready bucket should start reducing at once.  B200_COMM_PRIORITY=0
        # lets the comm / optimizer CTAs yield to the conv kernels instead.
        self.comm_stream = torch.cuda.Stream(device=dev, priority=int(os.environ.get("B200_COMM_PRIORITY", "-1")))
        # Second side stream for the optimizer: the reduction of bucket k+1 (NVLink-bound) runs while
        # bucket k is being updated (HBM-bound).  On ONE stream the two add up -- at 2 ranks 24 x
        # (all-reduce + Adam) is 5.0 ms of serial work against 4.6 ms of backward, i.e. the comm stream,
        # not the convolutions, ends the step (profiles/r2_timeline_n2_single_comm_stream.txt).
        self.opt_stream = torch.cuda.Stream(device=dev, priority=int(os.environ.get("B200_COMM_PRIORITY", "-1"))) \
            if os.environ.get("B200_SPLIT_COMM", "1") == "1" else self.comm_stream
        if self.world > 1 and not self.use_nccl:
            from ..parallel import topology
            from ..parallel.symm import SymmetricArena

            layout = topology.detect_layout()
            self.comm_mode = layout.mode()
            if self.comm_mode == "flat" and os.environ.get("B200_FORCE_HIERARCHICAL", "0") == "1":
                self.comm_mode = "hierarchical"          # test hook: one node, both stages exercised
            node_group = None
            if self.comm_mode == "hierarchical":
                node_group, self.cross_group = topology.make_hierarchy_groups(layout)
            if self.comm_mode == "nccl":
                # one GPU per host (the reference's deployment) or uneven hosts: no peer memory to fuse over
                self.use_nccl = True
                if self.rank == 0:
                    print(f"[b200] {layout.n_nodes} hosts x {layout.local_size} GPU: gradient all-reduce "
                          f"falls back to NCCL (--allreduce {allreduce} needs an NVLink domain)", flush=True)
            else:
                wdt = BF16 if wire_dtype == "bf16" else F32
                self.arena = SymmetricArena(n, dev, wire_dtype=wdt, group=node_group)
        if self.world > 1:
            self._broadcast_params()
        # buckets holding only FC weight gradients are written with plain stores by the wgrad GEMM:
        # the optimizer does not need to re-zero them (conv / bias gradients accumulate with red.add)
        fc_w = {f.name + ".weight" for f in spec.fcs}
        self._bucket_store_only = [all(t in fc_w for t in bk.tensors) for bk in self.plan.buckets]
        # FC weights that span whole buckets by themselves can be produced directly as bf16 on the wire
        self._prepacked = set()
        self._bucket_prepacked = [False] * len(self.plan.buckets)
        if self.arena is not None and self.arena.wire_dtype == BF16:
            for f in spec.fcs:
                name = f.name + ".weight"
                ids = self.plan.bucket_of(name)
                if ids and all(self.plan.buckets[i].tensors == (name,) for i in ids):
                    self._prepacked.add(name)
                    for i in ids:
                        self._bucket_prepacked[i] = True

        # EXPERIMENTAL (--zero1): buckets that take a reduce-scatter algorithm run reduce-scatter +
        # Adam on the owned cells + all-gather of the new bf16 weights as ONE kernel; the fp32 master
        # and moments of a cell are then only current on its owner (prepare_export gathers them).
        # zero1: True / False, or "auto" = on from 4 ranks up (measured, profiles/r2_*: at 8 ranks it
        # takes the step from 7.35 to 6.83 ms; at 2 ranks every rank still owns half of the optimizer
        # state and the fused kernel's thin CTAs are slower than all-reduce + replicated Adam).
        z1_ok = self.arena is not None and self.arena.wire_dtype == BF16 and self.cross_group is None \
            and optimizer == "adam"
        if zero1 == "auto":
            zero1 = z1_ok and self.world >= 4
        self.zero1 = bool(zero1) and z1_ok
        if zero1 and not self.zero1 and self.world > 1:
            raise ValueError("--zero1 needs Adam, the bf16 wire and all ranks in one NVLink domain")
        self._zero1_buckets: set = set()
        # the fused ZeRO-1 kernel keeps less in flight per CTA (128 threads, one switch reduction +
        # 96 bytes of optimizer state each): more, equally thin, CTAs
        self.zero1_ctas = int(os.environ.get("B200_ZERO1_CTAS", "128"))

        self._build_buffers()

    # ============================================================================ parameters
    def _view(self, arena: torch.Tensor, name: str) -> torch.Tensor:
        off = self.plan.offsets[name]
        return arena[off:off + self.plan.numels[name]].view(L.native_shape(self.spec, name))

    def import_state(self, state: Dict[str, torch.Tensor], sync: bool = True) -> None:
        """Load torch-layout tensors (checkpoint / oracle state dict) into the arenas."""
        with torch.no_grad():
            for name in self.spec.param_names:
                t = state[name].detach().to(self.device, F32)
                self._view(self.p32, name).copy_(L.to_native(self.spec, name, t))
            ops.require().cast_to_bf16(self.p32, self.w16)
        if sync and self.world > 1:
            self._broadcast_params()

    def prepare_export(self) -> None:
        """Collective.  Under --zero1 every rank holds the current fp32 master / moments only for the
        cells it owns: sum the owned pieces over ranks so that every rank has the full state again."""
        if not self._zero1_buckets:
            return
        self.sync()
        for arena in (self.p32, self.m32, self.v32):
            for bi in sorted(self._zero1_buckets):
                bk = self.plan.buckets[bi]
                tmp = torch.zeros(bk.end - bk.start, dtype=F32, device=self.device)
                for a, b in self.arena.owned_ranges(bk.start, bk.end - bk.start, self.zero1_ctas):
                    tmp[a - bk.start:b - bk.start] = arena[a:b]
                dist.all_reduce(tmp)
                arena[bk.start:bk.end] = tmp
        torch.cuda.synchronize(self.device)

    def export_state(self) -> Dict[str, torch.Tensor]:
        return {name: L.to_torch(self.spec, name, self._view(self.p32, name)).detach().cpu()
                for name in self.spec.param_names}

    def export_optimizer_state(self) -> dict:
        st = {"name": self.opt_name, "step": self.step_count, "lr": self.lr}
        if self.opt_name == "adam":
            st["exp_avg"] = {n: L.to_torch(self.spec, n, self._view(self.m32, n)).cpu() for n in self.spec.param_names}
            st["exp_avg_sq"] = {n: L.to_torch(self.spec, n, self._view(self.v32, n)).cpu() for n in self.spec.param_names}
        else:
            st["momentum_buffer"] = {n: L.to_torch(self.spec, n, self._view(self.m32, n)).cpu()
                                     for n in self.spec.param_names}
        return st

    def import_optimizer_state(self, st: dict) -> None:
        self.step_count = int(st.get("step", 0))
        if "lr" in st:                       # a run resumed behind an LR-step boundary keeps the decayed rate
            self.lr = float(st["lr"])
        with torch.no_grad():
            for key, arena in (("exp_avg", self.m32), ("exp_avg_sq", self.v32), ("momentum_buffer", self.m32)):
                if key in st and arena is not None:
                    for n, t in st[key].items():
                        self._view(arena, n).copy_(L.to_native(self.spec, n, t.to(self.device, F32)))

    def set_lr(self, lr: float) -> None:
        self.lr = lr

    def _broadcast_params(self) -> None:
        """DDP's constructor sync (distributedVggf.py:225): everybody adopts rank 0's weights."""
        if self.arena is not None and self.cross_group is None:
            self.arena.broadcast_(self.p32, root=0)
        else:
            dist.broadcast(self.p32, src=0)
        ops.require().cast_to_bf16(self.p32, self.w16)
        torch.cuda.synchronize(self.device)

    # ============================================================================== buffers
    def _build_buffers(self) -> None:
        spec, B, dev, HW = self.spec, self.B, self.device, self.HW
        # double-buffered device-side input staging, filled by a copy stream
        self.copy_stream = torch.cuda.Stream(device=dev)
        self._in_slot = 0
        self._src_u8 = [None, None]
        self._params_dev = [torch.zeros(B, 8, dtype=F32, device=dev) for _ in range(2)]
        self._labels_dev = [torch.zeros(B, dtype=torch.int64, device=dev) for _ in range(2)]
        self._in_free = [None, None]           # event: the step that used this slot is done with it
        self.labels_dev = self._labels_dev[0]
        # first conv: fused kernels build the im2col operand in shared memory from NHWC4 pixels
        # (csrc/include/conv0.cuh); the explicit im2col matrix only exists for other widths
        self.fused_conv0 = spec.convs[0].cout == 64
        self.col0 = None if self.fused_conv0 else torch.empty(B * HW * HW, L.CONV0_K, dtype=BF16, device=dev)
        self.x_nhwc = torch.empty(B, HW, HW, 4, dtype=BF16, device=dev)
        self.acts: List[torch.Tensor] = []      # post-ReLU conv outputs
        self.pools: List[Optional[torch.Tensor]] = []
        # EXPERIMENTAL, off unless B200_FUSE_POOL=1: the pooled layers' conv epilogue pools in
        # registers and records argmax / ReLU bit masks; the un-pooled activation is never written
        # (acts[i] of such a layer then holds stale data) and backward is an "unpool" kernel.
        want_fused_pool = os.environ.get("B200_FUSE_POOL", "1") == "1"
        self.pool_masks: List[Optional[torch.Tensor]] = []
        h = HW
        for i, c in enumerate(spec.convs):
            self.acts.append(torch.empty(B, h, h, c.cout, dtype=BF16, device=dev))
            fused = (want_fused_pool and c.pool_after and i > 0 and c.cout % 32 == 0
                     and bool(ops.require().conv_pool_fusable(B, h, h)))
            self.pool_masks.append(ops.pool_mask_like(B, h, h, c.cout, dev) if fused else None)
            if c.pool_after:
                h //= 2
                self.pools.append(torch.empty(B, h, h, c.cout, dtype=BF16, device=dev))
            else:
                self.pools.append(None)
        self.feat_hw = h
        cl = spec.convs[-1].cout
        ph = spec.pooled_hw
        assert spec.fcs[0].fin == cl * ph * ph, "classifier.0 must consume the pooled feature map"
        self.avg = None if h == ph else torch.empty(B, ph, ph, cl, dtype=BF16, device=dev)
        # FC stack
        self.fc_acc = [torch.zeros(B, f.fout, dtype=F32, device=dev) for f in spec.fcs]
        self.fc_y = [torch.empty(B, f.fout, dtype=BF16, device=dev) if i < len(spec.fcs) - 1 else None
                     for i, f in enumerate(spec.fcs)]
        self.fc_dacc = [torch.zeros(B, f.fin, dtype=F32, device=dev) for f in spec.fcs]
        self.fc_dz = [torch.zeros(B, _round_up(f.fout, 8), dtype=BF16, device=dev) for f in spec.fcs]
        self.logits = torch.zeros(B, spec.num_classes, dtype=F32, device=dev)
        self.loss_buf = torch.zeros(1, dtype=F32, device=dev)
        self.scratch_meter = torch.zeros(4, dtype=F32, device=dev)
        # K-FUN2+CE: last Linear + cross-entropy + metrics + that layer's backward as ONE launch when the
        # head is small (the funnel: 512 -> C <= 8); VGG-16/1000 keeps the GEMM path.  B200_FUSED_HEAD=0: off.
        lastf = spec.fcs[-1]
        self.fused_head = (os.environ.get("B200_FUSED_HEAD", "1") == "1" and len(spec.fcs) >= 2
                           and bool(ops.require().head_ce_supported(B, lastf.fout, lastf.fin)))
        self._head_in = None
        # backward activations-gradients: two ping-pong buffers big enough for the largest map
        biggest = max(a.numel() for a in self.acts)
        self.dbuf = [torch.empty(biggest, dtype=BF16, device=dev) for _ in range(2)]
        self.dfeat = torch.empty(B, ph, ph, cl, dtype=BF16, device=dev)

    # ================================================================================ inputs
    def _stage_input(self, batch) -> int:
        """Copy a batch to the device and produce the first conv's im2col rows.  Returns b."""
        C = ops.require()
        HW = self.HW
        j = self._in_slot
        self._in_slot ^= 1
        self.labels_dev = self._labels_dev[j]
        cur = torch.cuda.current_stream(self.device)
        if isinstance(batch, FusedBatch):
            b = int(batch.labels.shape[0])
            src = batch.images_u8
            if self._src_u8[j] is None or self._src_u8[j].shape[1:] != src.shape[1:]:
                self._src_u8[j] = torch.empty((self.B,) + tuple(src.shape[1:]), dtype=torch.uint8,
                                              device=self.device)
            with torch.cuda.stream(self.copy_stream):
                if self._in_free[j] is not None:
                    self.copy_stream.wait_event(self._in_free[j])
                self._src_u8[j][:b].copy_(src, non_blocking=True)
                self._params_dev[j][:b].copy_(batch.params, non_blocking=True)
                self._labels_dev[j][:b].copy_(batch.labels, non_blocking=True)
                h2d = torch.cuda.Event()
                h2d.record(self.copy_stream)
            if batch.state is not None:
                batch.state["event"] = h2d          # the loader may refill the pinned slot after this
            cur.wait_event(h2d)
            # two passes: transform once per pixel into 8-byte NHWC4 pixels (25 MB at B=64, L2
            # resident), then expand to the first conv's im2col rows with coalesced 16-byte stores
            ops.augment(self._src_u8[j][:b], self._params_dev[j][:b], self.x_nhwc[:b], batch.resized_hw,
                        out_hw=HW, mode="nhwc", pad=4)
            if not self.fused_conv0:
                C.im2col_c3(self.x_nhwc[:b], self.col0[:b * HW * HW], L.CONV0_K)
            return b
        x, y = batch
        b = int(y.shape[0])
        x = x.to(self.device, F32, non_blocking=True).contiguous()
        assert x.shape[2] == HW and x.shape[3] == HW, "input size differs from the engine's input_hw"
        self.labels_dev[:b].copy_(y, non_blocking=True)
        C.nchw_to_nhwc(x, self.x_nhwc[:b])
        if not self.fused_conv0:
            C.im2col_c3(self.x_nhwc[:b], self.col0[:b * HW * HW], L.CONV0_K)
        return b

    def _release_input(self) -> None:
        """The augment kernel and the loss have consumed the staged batch: its slot may be refilled."""
        ev = torch.cuda.Event()
        ev.record()
        self._in_free[self._in_slot ^ 1] = ev

    # =============================================================================== forward
    def _w(self, name: str) -> torch.Tensor:
        return self._view(self.w16, name + ".weight")

    def _b(self, name: str) -> torch.Tensor:
        return self._view(self.p32, name + ".bias")

    def _drop_key(self, layer: int) -> int:
        """Philox key of a dropout layer.  Every replica draws its OWN mask, like the reference's
        per-process RNG (distributedVggf.py:55; one process per rank) -- identical masks on all ranks
        would correlate the replicas' gradient noise.  ``same_dropout_all_ranks`` is for parity tests."""
        r = 0 if self.same_dropout_all_ranks else self.rank
        return self.seed * 1000003 + layer + 7919 * r

    @staticmethod
    def _ksplit(m_tiles: int, k_iters: int, target: int = 256) -> int:
        return max(1, min(k_iters, target // max(m_tiles, 1)))

    def _forward(self, b: int, train: bool) -> None:
        C = ops.require()
        spec, HW = self.spec, self.HW
        c0 = spec.convs[0]
        M0 = b * HW * HW
        # conv0 (K = 27): im2col GEMM with fused bias + ReLU -> bf16 NHWC
        if self.fused_conv0:
            C.conv0_fprop(self.x_nhwc[:b], self._w(c0.name), self._b(c0.name), self.acts[0][:b])
        else:
            ops.gemm(self.col0[:M0], self._w(c0.name), self.acts[0][:b].view(M0, c0.cout), M=M0, N=c0.cout,
                     K=L.CONV0_K, epi="bf16_bias_relu", bias=self._b(c0.name))
        x = self.acts[0][:b]
        if c0.pool_after:
            x = ops.maxpool2x2(x, out=self.pools[0][:b])
        for i, c in enumerate(spec.convs[1:], start=1):
            if self.pool_masks[i] is not None:           # experimental fused conv + ReLU + pool
                x, _ = ops.conv3x3_fprop_pool(x, self._w(c.name), self._b(c.name), out=self.pools[i][:b],
                                              mask=self.pool_masks[i][:b])
                self._mark("fwd " + c.name)
                continue
            y = self.acts[i][:b]
            C.conv_fprop(x, self._w(c.name), self._b(c.name), y, True, 0)
            x = y
            if c.pool_after:
                x = ops.maxpool2x2(y, out=self.pools[i][:b])
            self._mark("fwd " + c.name)
        if self.avg is not None:
            x = ops.adaptive_avgpool(x, spec.pooled_hw, spec.pooled_hw, out=self.avg[:b])
        self.feat = x
        h = x.reshape(b, -1)
        last = len(spec.fcs) - 1
        for i, f in enumerate(spec.fcs):
            if i == last and self.fused_head:
                self._head_in = h            # logits, loss and this layer's backward: _head()
                break
            m_tiles = (f.fout + 127) // 128
            ks = self._ksplit(m_tiles, (f.fin + 63) // 64)
            acc = self.fc_acc[i][:b]
            ops.gemm(self._w(f.name), h, acc, M=f.fout, N=b, K=f.fin, epi="f32_atomic_t", ksplit=ks, ldo=f.fout)
            if i == last:
                ops.fc_bias_act(acc, self._b(f.name), None, self.logits[:b], B=b, N=f.fout, relu=False)
            else:
                p = f.dropout if (train and self.train_dropout) else 0.0
                ops.fc_bias_act(acc, self._b(f.name), self.fc_y[i][:b], None, B=b, N=f.fout, relu=f.relu,
                                drop_p=p, seed=self._drop_key(i), offset=self.step_count)
                h = self.fc_y[i][:b]

    def _loss(self, b: int, train: bool) -> None:
        """Loss + metrics (+ dlogits when training) from the forward state."""
        meter = self.meter.buf if self.meter is not None else self.scratch_meter
        if not self.fused_head:
            ld = self.fc_dz[-1].shape[1]
            ops.cross_entropy(self.logits[:b], self.labels_dev[:b], self.fc_dz[-1][:b] if train else None,
                              ld if train else 0, meter, self.loss_buf,
                              class_weights=self.class_weights if train else None)   # like the reference's loops
            return
        f, prev = self.spec.fcs[-1], self.spec.fcs[-2]
        C = ops.require()
        if train:
            p = prev.dropout if self.train_dropout else 0.0
            C.head_ce(self._head_in, self._w(f.name), self._b(f.name), self.labels_dev[:b], self.logits[:b],
                      self.fc_dz[-1][:b], self.fc_dz[-1].shape[1], self._grad(f.name + ".weight"),
                      self._grad(f.name + ".bias"), self.fc_dz[-2][:b], 1.0 / (1.0 - p), bool(prev.relu), meter,
                      self.loss_buf, self.class_weights)
        else:
            C.head_ce(self._head_in, self._w(f.name), self._b(f.name), self.labels_dev[:b], self.logits[:b],
                      None, 0, None, None, None, 1.0, False, meter, self.loss_buf, None)

    # ============================================================================== backward
    def _grad(self, name: str) -> torch.Tensor:
        return self._view(self.g32, name)

    def _backward(self, b: int) -> None:
        C = ops.require()
        spec = self.spec
        fcs, convs = spec.fcs, spec.convs
        last = len(fcs) - 1
        # ---- FC stack ----------------------------------------------------------------------
        for i in range(last, -1, -1):
            f = fcs[i]
            if i == last and self.fused_head:     # produced by _loss(): gradients are already in the arena
                self._bucket_done(f.name + ".bias")
                self._bucket_done(f.name + ".weight")
                self._mark("bwd " + f.name)
                continue
            ld = self.fc_dz[i].shape[1]
            dz = self.fc_dz[i][:b]                                   # [b][ld] bf16, cols >= fout are 0
            x_in = self.fc_y[i - 1][:b] if i > 0 else self.feat.reshape(b, -1)
            ops.bias_grad(dz, self._grad(f.name + ".bias"), b, f.fout, ld=ld)
            self._bucket_done(f.name + ".bias")
            # wgrad: dW[fout][fin] = dz^T x_in   (both operands MN-major: stored [K=b][*])
            if (f.name + ".weight") in self._prepacked:
                # big FC weights own whole buckets: write bf16(dW / ws) straight into the symmetric
                # wire buffer -- no fp32 gradient, no pack pass before the reduction
                off = self.plan.offsets[f.name + ".weight"]
                wire_w = self.arena.wire[off:off + f.fout * f.fin].view(f.fout, f.fin)
                ops.gemm(dz, x_in, wire_w, M=f.fout, N=f.fin, K=b, a_mn=True, b_mn=True, epi="bf16_store",
                         ldo=f.fin, alpha=1.0 / self.world)
            else:
                ops.gemm(dz, x_in, self._grad(f.name + ".weight"), M=f.fout, N=f.fin, K=b, a_mn=True,
                         b_mn=True, epi="f32_store", ldo=f.fin)
            # dgrad: dX[b][fin] = dz W ; swap-AB with W read MN-major (stored [K=fout][M=fin])
            m_tiles = (f.fin + 127) // 128
            ks = self._ksplit(m_tiles, (f.fout + 63) // 64)
            dacc = self.fc_dacc[i][:b]
            ops.gemm(self._w(f.name), dz, dacc, M=f.fin, N=b, K=f.fout, a_mn=True,
                     epi="f32_atomic_t" if ks > 1 else "f32_store_t", ksplit=ks, ldo=f.fin)
            # only now may the optimizer touch W_i: the dgrad above still reads the old weights
            self._bucket_done(f.name + ".weight")
            if i > 0:
                prev = fcs[i - 1]
                ops.fc_grad_act(dacc, self.fc_y[i - 1][:b], self.fc_dz[i - 1][:b], B=b, N=f.fin,
                                relu=prev.relu, drop_p=prev.dropout if self.train_dropout else 0.0)
            else:
                ops.fc_grad_act(dacc, None, self.dfeat[:b].view(b, -1), B=b, N=f.fin, relu=False)
            self._mark("bwd " + f.name)
        # ---- feature map gradient ------------------------------------------------------------
        g = self.dfeat[:b]
        if self.avg is not None:
            g = ops.adaptive_avgpool_bwd(g, self.feat_hw, self.feat_hw,
                                         out=self.dbuf[1][:b * self.feat_hw ** 2 * convs[-1].cout]
                                         .view(b, self.feat_hw, self.feat_hw, convs[-1].cout))
        # ---- conv stack ----------------------------------------------------------------------
        # invariant at loop entry: if layer i pools, `g` is dP_i (grad wrt pooled output);
        # otherwise `g` is dZ_i (grad wrt the pre-ReLU output, mask already applied).
        pp = 0
        bias_fused = False      # was this layer's bias gradient already produced by the kernel that made dz?
        for i in range(len(convs) - 1, -1, -1):
            c = convs[i]
            y = self.acts[i][:b]
            if c.pool_after:
                dz = self.dbuf[pp][:y.numel()].view_as(y)
                # ReLU + pool backward, bias gradient (column sum of dz) fused
                if self.pool_masks[i] is not None:
                    ops.unpool2x2(g, self.pool_masks[i][:b], out=dz, colsum=self._grad(c.name + ".bias"))
                else:
                    ops.maxpool2x2_relu_bwd(y, g, out=dz, colsum=self._grad(c.name + ".bias"))
                pp ^= 1
                bias_fused = True
            else:
                dz = g
            M = y.numel() // c.cout
            if not bias_fused:
                ops.bias_grad(dz.view(M, c.cout), self._grad(c.name + ".bias"), M, c.cout)
            self._bucket_done(c.name + ".bias")
            if i == 0:
                if self.fused_conv0:
                    C.conv0_wgrad(dz, self.x_nhwc[:b], self._grad(c.name + ".weight"))
                else:
                    ks = max(1, min(M // 64, 296))
                    ops.gemm(dz.view(M, c.cout), self.col0[:M], self._grad(c.name + ".weight"), M=c.cout,
                             N=L.CONV0_K, K=M, a_mn=True, b_mn=True, epi="f32_atomic", ksplit=ks, ldo=L.CONV0_K)
                self._bucket_done(c.name + ".weight")
                self._mark("bwd " + c.name)
                break
            prev = convs[i - 1]
            x_in = self.pools[i - 1][:b] if prev.pool_after else self.acts[i - 1][:b]
            C.conv_wgrad(dz, x_in, self._grad(c.name + ".weight"), 1.0, 0, 0)
            dx = self.dbuf[pp][:x_in.numel()].view_as(x_in)
            # ReLU backward of layer i-1 is fused here unless a pool sits in between; when it is,
            # dx IS dZ_{i-1} and its column sum (layer i-1's bias gradient) comes for free
            fuse = not prev.pool_after
            C.conv_dgrad(dz, self._w(c.name), x_in if fuse else None, dx,
                         self._grad(prev.name + ".bias") if fuse else None, 0)
            bias_fused = fuse
            self._bucket_done(c.name + ".weight")      # after dgrad: it reads the pre-update weights
            self._mark("bwd " + c.name)
            pp ^= 1
            g = dx

    # ================================================================= all-reduce + optimizer
    def _begin_step(self) -> None:
        self._missing = [len(bk.tensors) for bk in self.plan.buckets]
        self._tensor_buckets = getattr(self, "_tensor_buckets", None) or \
            {n: self.plan.bucket_of(n) for n in self.plan.order}

    def _bucket_done(self, name: str) -> None:
        """Called right after the kernels producing ``name``'s gradient were enqueued."""
        for bi in self._tensor_buckets[name]:
            self._missing[bi] -= 1
            if self._missing[bi] == 0:
                self._reduce_and_update(bi)

    def _apply_update(self, s: int, e: int, g16: Optional[torch.Tensor], zero: bool = True) -> None:
        g32 = None if g16 is not None else self.g32[s:e]
        z = self.g32[s:e] if zero else None      # ranges written with plain stores need no re-zeroing
        if self.opt_name == "adam":
            ops.adam_step(self.p32[s:e], self.m32[s:e], self.v32[s:e], g32=g32, g16=g16, shadow=self.w16[s:e],
                          lr=self.lr, beta1=self.beta1, beta2=self.beta2, eps=self.eps,
                          weight_decay=self.weight_decay, step=self.step_count, zero=z)
        else:
            ops.sgd_step(self.p32[s:e], self.m32[s:e], g32=g32, g16=g16, shadow=self.w16[s:e], lr=self.lr,
                         momentum=self.momentum, weight_decay=self.weight_decay,
                         first=(self.step_count == 1), zero=z)

    def _reduce_and_update(self, bi: int) -> None:
        bk = self.plan.buckets[bi]
        s, e = bk.start, bk.end
        if self.world == 1:
            if self.apply_updates:
                # The optimizer is HBM-bound, the convolutions still to come are tensor-bound: run the
                # update of a finished bucket on the side stream so it hides under the rest of backward
                # (FC-1/FC-2 are 88 % of the parameters and finish first).
                ev = torch.cuda.Event()
                ev.record()
                self.comm_stream.wait_event(ev)
                with torch.cuda.stream(self.comm_stream):
                    self._mark("opt%d start" % bi, "comm")
                    self._apply_update(s, e, None, zero=not self._bucket_store_only[bi])
                    self._mark("opt%d end" % bi, "comm")
            return
        ev = torch.cuda.Event()
        ev.record()
        self.comm_stream.wait_event(ev)
        with torch.cuda.stream(self.comm_stream):
            if self.use_nccl:                    # library baseline (comparison only)
                self.g32[s:e].mul_(1.0 / self.world)
                dist.all_reduce(self.g32[s:e])
                if self.apply_updates:
                    self._apply_update(s, e, None)
                return
            algo = self.arena.pick_algo(e - s, self.ar_algo)
            prepacked = self._bucket_prepacked[bi]
            if prepacked and algo == "oneshot":
                algo = "twoshot"                 # one-shot cannot leave its result on the wire
            # ZeRO-1 only for buckets that hold nothing but one big FC weight (88 % of the parameters):
            # the forward pass reads those through the all-gathered bf16 shadow.  Buckets with biases
            # stay replicated -- biases are read from the fp32 master, which under ZeRO-1 is current on
            # the owner rank only.
            if self.zero1 and self.apply_updates and prepacked and algo != "oneshot":
                self._mark("z1_%d start" % bi, "comm")
                zev0 = None
                if self._comm_events is not None:
                    zev0 = torch.cuda.Event(enable_timing=True)
                    zev0.record()
                self.arena.zero1_step(None if prepacked else self.g32, self.p32, self.m32, self.v32, self.w16,
                                      s, e - s, algo=algo, slot=bi % self.arena.slots, max_ctas=self.zero1_ctas,
                                      inv_world=1.0 / self.world, lr=self.lr, beta1=self.beta1, beta2=self.beta2,
                                      eps=self.eps, weight_decay=self.weight_decay, step=self.step_count)
                self._zero1_buckets.add(bi)
                if zev0 is not None:             # reduce-scatter + optimizer + all-gather: same wire bytes
                    zev1 = torch.cuda.Event(enable_timing=True)
                    zev1.record()
                    self._comm_events.append((zev0, zev1, (e - s) * self.arena.wire.element_size()))
                self._mark("z1_%d end" % bi, "comm")
                return
            to_f32 = (algo == "oneshot" or self.unpack_fp32 or self.arena.wire_dtype == F32
                      or not self.apply_updates)
            self._mark("ar%d start" % bi, "comm")
            ev0 = None
            if self._comm_events is not None:    # observability: device time of every reduction
                ev0 = torch.cuda.Event(enable_timing=True)
                ev0.record()
            self.arena.allreduce(None if prepacked else self.g32, self.g32 if to_f32 else None, s, e - s,
                                 algo=algo, slot=bi % self.arena.slots, max_ctas=self.comm_ctas,
                                 inv_world=1.0 / self.world)
            if ev0 is not None:
                ev1 = torch.cuda.Event(enable_timing=True)
                ev1.record()
                self._comm_events.append((ev0, ev1, (e - s) * self.arena.wire.element_size()))
            if self.cross_group is not None:     # node sums (already x 1/world) -> job sum, over NCCL
                dist.all_reduce(self.g32[s:e] if to_f32 else self.arena.wire[s:e], group=self.cross_group)
            self._mark("ar%d end" % bi, "comm")
            if self.apply_updates:
                if self.opt_stream is not self.comm_stream:
                    red = torch.cuda.Event()
                    red.record()                      # on the comm stream: bucket bi is reduced
                    self.opt_stream.wait_event(red)
                with torch.cuda.stream(self.opt_stream):
                    self._apply_update(s, e, None if to_f32 else self.arena.wire[s:e],
                                       zero=not (prepacked or self._bucket_store_only[bi]))
                    self._mark("opt%d end" % bi, "opt")

    def _end_step(self) -> None:
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.comm_stream)
        if self.opt_stream is not self.comm_stream:
            cur.wait_stream(self.opt_stream)

    # ================================================================================ public
    def set_meter(self, meter: Optional[DeviceMeter]) -> None:
        self.meter = meter

    def set_class_weights(self, weights: Optional[torch.Tensor]) -> None:
        """Per-class loss weights (the reference's optional weighted cross-entropy)."""
        self.class_weights = None if weights is None else weights.to(self.device, F32).contiguous()

    def train_step(self, batch) -> torch.Tensor:
        """One optimisation step.  Returns the device scalar holding this batch's mean loss."""
        self.step_count += 1
        t = self.timer
        if self.nvtx:
            torch.cuda.nvtx.range_push("step")
        self._mark("step start")
        t.start("input")
        b = self._stage_input(batch)
        t.stop("input")
        self._mark("input")
        t.start("forward")
        self._forward(b, train=True)
        self._loss(b, train=True)
        self._release_input()
        self._mark("loss")
        t.stop("forward")
        t.start("backward+reduce+update")
        self._begin_step()
        self._backward(b)
        self._end_step()
        self._mark("step end (joined comm)")
        t.stop("backward+reduce+update")
        if self.nvtx:
            torch.cuda.nvtx.range_pop()
        return self.loss_buf

    def eval_step(self, batch) -> torch.Tensor:
        b = self._stage_input(batch)
        self._forward(b, train=False)
        self._loss(b, train=False)
        self._release_input()
        return self.loss_buf

    def forward_logits(self, batch) -> torch.Tensor:
        b = self._stage_input(batch)
        self._forward(b, train=False)
        if self.fused_head:          # the logits come out of the fused head kernel
            keep = self.meter
            self.meter = None
            self._loss(b, train=False)
            self.meter = keep
        self._release_input()
        return self.logits[:b].clone()

    def sync(self) -> None:
        torch.cuda.synchronize(self.device)

    # ---- step timeline (SURVEY 5.1 tracing; there is no nsys on the box) ----------------------------
    def timeline(self, on: bool) -> None:
        """Record a CUDA event after every layer's kernels on the compute stream and around every
        all-reduce / optimizer launch on the comm stream.  ``timeline_report()`` turns them into
        milliseconds since the start of the step: which bucket's reduction ran under which layer's
        backward, and what (if anything) is exposed after the last wgrad."""
        self._tl = [] if on else None

    def _mark(self, name: str, lane: str = "compute") -> None:
        if self._tl is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()                      # on the current stream
            self._tl.append((name, lane, ev))

    def timeline_report(self) -> list:
        """[(name, lane, ms since the step's first mark)] for everything recorded since timeline(True)."""
        if not self._tl:
            return []
        self.sync()
        t0 = self._tl[0][2]
        out = [(n, lane, round(t0.elapsed_time(ev), 4)) for n, lane, ev in self._tl]
        self._tl = []
        return out

    # ---- gradient all-reduce observability (GB/s and roofline fraction) -----------------
    def comm_timing(self, on: bool) -> None:
        """Record CUDA events around every fused all-reduce launch (comm stream).  The times include
        whatever the kernel waits for (peers arriving late) and run under backward's kernels -- it is
        the rate the training step actually sees, not an isolated micro-benchmark."""
        self._comm_events = [] if on else None

    def comm_report(self, steps: int) -> Optional[dict]:
        """Device time and bus bandwidth of the gradient all-reduces of the last ``steps`` steps:
        bus GB/s = 2(ws-1)/ws * wire bytes / time (NCCL's convention), against the measured 770 GB/s
        peer-copy rate and the nominal 900 GB/s per direction."""
        if not self._comm_events or self.world == 1:
            return None
        self.sync()
        ms = sum(a.elapsed_time(b) for a, b, _ in self._comm_events)
        nbytes = sum(n for _, _, n in self._comm_events)
        self._comm_events = []
        bus = 2.0 * (self.world - 1) / self.world * nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"wire_MB_per_step": round(nbytes / steps / 1e6, 2), "ms_per_step": round(ms / steps, 4),
                "launches_per_step": len(self.plan.buckets), "bus_GBs": round(bus, 1),
                "frac_of_770_measured": round(bus / 770.0, 3), "frac_of_900_nominal": round(bus / 900.0, 3),
                "comm_ctas": self.comm_ctas, "overlapped_with_backward": True,
                "includes_fused_optimizer": bool(self._zero1_buckets)}

    def phase_times(self) -> Dict[str, float]:
        return self.timer.collect()

    # torch-module-like conveniences used by Trainer / checkpoint
    def train(self) -> None:
        pass

    def eval(self) -> None:
        pass
