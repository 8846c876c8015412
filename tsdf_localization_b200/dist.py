"""One-process-per-GPU driver of the sensor update: particles sharded, TSDF map replicated.

Rank r evaluates the contiguous particle slice [r*ceil(N/R), ...) against the full scan; the un-normalised weights are
all-gathered (NCCL over NVLink; N fp32); every rank then normalises and scans the GLOBAL weight vector redundantly —
same kernels, same fixed summation order, so all ranks hold bit-identical CDFs and the result is independent of R —
draws only ITS contiguous slice of output slots, and the resampled particles are all-gathered (7*N fp32) so the next
update again finds the full particle set on every GPU. One host sync per update (status read-back: n_out, zero-sum).

The reference has no multi-GPU path at all (single process, default stream); this is the north star's
item (4). ``torch`` is plumbing here: device buffers, the stream, ``torch.distributed`` collectives. The stage
implementation is injectable so the slicing / padding / collective logic is testable on CPU with gloo (tests/ plug the
CPU oracle in; the product class ``GpuStages`` only ever calls libtsdfloc.so).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import capi


def shard(n: int, world: int, rank: int) -> Tuple[int, int, int]:
    """(chunk, first, count): contiguous slices of ceil(n / world); trailing ranks may get a short or empty slice."""
    chunk = (n + world - 1) // world
    first = min(rank * chunk, n)
    return chunk, first, min(chunk, n - first)


def output_capacity(n: int, world: int) -> int:
    """Output slots reserved for resampling n particles. The reference recurrence emits exactly n particles when n is a
    power of two (1/n exact in fp32) and drifts by up to a few per cent otherwise."""
    cap = n if (n & (n - 1)) == 0 else n + n // 8 + 64
    return ((cap + world - 1) // world) * world


class GpuStages:
    """The device-pointer stage calls of include/tsdfloc.h on torch CUDA tensors and the current torch stream."""

    def __init__(self, evaluator):
        self.ev = evaluator
        self.lib = capi.load_library()
        self.ctx = evaluator.ctx

    @staticmethod
    def _stream() -> C.c_void_p:
        # NULL means "the ctx's own stream" in the C ABI; torch's default stream has handle 0, which is CUDA's legacy
        # default stream -> pass cudaStreamLegacy (0x1) so the kernels are ordered with torch's and NCCL's work on it.
        h = torch.cuda.current_stream().cuda_stream
        return C.c_void_p(h if h else 1)

    def set_scan(self, d_points: torch.Tensor) -> None:
        assert d_points.is_cuda and d_points.dtype == torch.float32 and d_points.is_contiguous()
        capi.check(self.lib, self.ctx, self.lib.tsdfloc_set_scan_device(self.ctx, C.c_void_p(d_points.data_ptr()), d_points.shape[0],
                                                                       self._stream()))

    def eval(self, d_particles, n, first, count, tf, d_raw) -> None:
        tfc = (C.c_float * 16)(*[float(v) for v in tf])
        capi.check(self.lib, self.ctx, self.lib.tsdfloc_eval_device(self.ctx, C.c_void_p(d_particles.data_ptr()), n, first, count, tfc,
                                                                   C.c_void_p(d_raw.data_ptr()), self._stream()))

    def eval_peers(self, d_particles, n, first, count, tf, d_raw, peer_raw_ptrs) -> None:
        """eval() whose final stores also go into every peer's weight vector (fused all-gather over NVLink)."""
        tfc = (C.c_float * 16)(*[float(v) for v in tf])
        arr = (C.c_void_p * len(peer_raw_ptrs))(*[int(p) for p in peer_raw_ptrs])
        capi.check(self.lib, self.ctx, self.lib.tsdfloc_eval_device_peers(self.ctx, C.c_void_p(d_particles.data_ptr()), n, first, count, tfc,
                                                                         C.c_void_p(d_raw.data_ptr()), arr, len(peer_raw_ptrs), self._stream()))

    def draw_peers(self, d_particles, n, u0, first_out, count_out, d_out, peer_out_ptrs) -> None:
        """draw() whose output slice is also stored into every peer's particle buffer (peer_out_ptrs[r] = address of slot
        first_out in rank r's buffer)."""
        arr = (C.c_void_p * len(peer_out_ptrs))(*[int(p) for p in peer_out_ptrs])
        capi.check(self.lib, self.ctx, self.lib.tsdfloc_draw_device_peers(self.ctx, C.c_void_p(d_particles.data_ptr()), n, C.c_float(u0),
                                                                         first_out, count_out, C.c_void_p(d_out.data_ptr()), arr,
                                                                         len(peer_out_ptrs), None, self._stream()))

    def update_fused(self, d_points, d_particles, n, tf, u0, d_out, count_out, d_mean) -> None:
        """Single-GPU update in one call (tsdfloc_update_device): scan preparation + matrices, evaluation, normalisation, draw."""
        tfc = (C.c_float * 16)(*[float(v) for v in tf])
        capi.check(self.lib, self.ctx, self.lib.tsdfloc_update_device(
            self.ctx, C.c_void_p(d_points.data_ptr()), d_points.shape[0], C.c_void_p(d_particles.data_ptr()), n, tfc, C.c_float(u0),
            C.c_void_p(d_out.data_ptr()), count_out, C.c_void_p(d_mean.data_ptr()), self._stream()))

    def normalize(self, d_particles, n, d_raw, d_mean) -> None:
        capi.check(self.lib, self.ctx, self.lib.tsdfloc_normalize_device(self.ctx, C.c_void_p(d_particles.data_ptr()), n,
                                                                        C.c_void_p(d_raw.data_ptr()), C.c_void_p(d_mean.data_ptr()),
                                                                        self._stream()))

    def draw(self, d_particles, n, u0, first_out, count_out, d_out) -> None:
        capi.check(self.lib, self.ctx, self.lib.tsdfloc_draw_device(self.ctx, C.c_void_p(d_particles.data_ptr()), n, C.c_float(u0),
                                                                   first_out, count_out, C.c_void_p(d_out.data_ptr()), None,
                                                                   self._stream()))

    def motion(self, d_particles, n, mean, sigma, seed: int, sequence: int, d_draws=None) -> None:
        """In-place motion update of the (replicated) particle set: every rank applies the same Philox stream, keyed by the
        particle index, so the replicas stay bit-identical without any exchange."""
        m = (C.c_double * 6)(*[float(v) for v in mean])
        s = (C.c_double * 6)(*[float(v) for v in sigma])
        capi.check(self.lib, self.ctx, self.lib.tsdfloc_motion_update_device(
            self.ctx, C.c_void_p(d_particles.data_ptr()), n, m, s, C.c_void_p(d_draws.data_ptr()) if d_draws is not None else None,
            int(seed), int(sequence), self._stream()))

    def check(self) -> Tuple[int, float]:
        n_out, wsum = C.c_uint64(0), C.c_double(0.0)
        rc = self.lib.tsdfloc_check(self.ctx, C.byref(n_out), C.byref(wsum), self._stream())
        if rc == capi.E_NO_VALID_PARTICLE:
            raise RuntimeError("No particle is valid!")
        capi.check(self.lib, self.ctx, rc)
        return int(n_out.value), float(wsum.value)

    def last_eval_ms(self) -> float:
        ms = C.c_float(0.0)
        capi.check(self.lib, self.ctx, self.lib.tsdfloc_last_eval_ms(self.ctx, C.byref(ms)))
        return float(ms.value)

    def kernel_launches(self) -> int:
        return int(self.lib.tsdfloc_kernel_launches(self.ctx))


class ShardedSensorUpdate:
    """Full sensor update (evaluation + normalisation + systematic resampling) over ``world`` ranks.

    All ranks hold the full particle set ``particles[:n]`` (float32[cap, 7], Particle layout) and the scan; after
    ``step`` they all hold the full resampled set. With world == 1 no collective is issued.
    """

    def __init__(self, stages, world: int = 1, rank: int = 0, group=None, device: Optional[torch.device] = None,
                 max_particles: int = 0, all_gather=None, fused: Optional[bool] = None):
        """fused: True = the evaluation and draw kernels store their results straight into every peer's buffers (symmetric
        memory over NVLink; three signal barriers per update instead of two NCCL all-gathers); False = NCCL all-gathers;
        None = fused when the stages support it and symmetric memory can be set up on this device, else NCCL."""
        self.stages = stages
        self.world, self.rank, self.group = int(world), int(rank), group
        self.transport = "none" if self.world == 1 else "all_gather"
        self._want_fused = fused
        self._symm = None
        # all_gather(out_flat, in_flat): torch.distributed (NCCL on GPUs, gloo in the CPU tests) unless injected
        self._all_gather = all_gather if all_gather is not None else self._dist_all_gather
        self.device = device if device is not None else torch.device("cpu")
        self._cap = 0
        self._reserve(max_particles)

    def _dist_all_gather(self, out: torch.Tensor, inp: torch.Tensor) -> None:
        import torch.distributed as dist
        dist.all_gather_into_tensor(out, inp, group=self.group)

    def _reserve(self, n: int) -> None:
        if n <= self._cap:
            return
        W = self.world
        chunk = (n + W - 1) // W
        ocap = output_capacity(n, W)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.mean = torch.zeros(8, **f32)
        self._cap = n
        self._symm = None
        want = self._want_fused
        can = (W > 1 and self._all_gather == self._dist_all_gather and self.device.type == "cuda" and hasattr(self.stages, "eval_peers"))
        if want is not False and can:
            try:
                self._setup_symmetric(W * chunk, ocap)
                self.transport = "fused_p2p"
                return
            except Exception as e:  # noqa: BLE001
                if want:
                    raise
                import warnings
                warnings.warn(f"symmetric memory unavailable ({e}); using NCCL all-gathers")
        elif want and W > 1:
            raise RuntimeError("fused=True needs CUDA devices, torch.distributed collectives and peer-capable stages")
        self.raw = torch.zeros(W * chunk, **f32)              # un-normalised weights, slice r at [r*chunk, (r+1)*chunk)
        # resampled particles, slice r at [r*ochunk, (r+1)*ochunk); two copies used alternately, so that the set a step
        # returns (a view into one copy) can be fed straight back in as the next step's `particles`
        self._out2 = torch.zeros((2, ocap, 7), **f32)
        self.out = self._out2[0]
        self._out_tick = 0

    def _setup_symmetric(self, n_raw: int, ocap: int) -> None:
        """Weight vector and output particle buffer in symmetric memory: every rank maps every peer's copy."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = self.group if self.group is not None else dist.group.WORLD
        # two copies of each buffer, used alternately: a rank may start storing update k+1 into its peers while they still
        # read update k (the barrier inside update k+1 orders it against update k-1, the last user of the same copy)
        raw = symm_mem.empty((2, n_raw), dtype=torch.float32, device=self.device)
        out = symm_mem.empty((2, ocap, 7), dtype=torch.float32, device=self.device)
        raw.zero_()
        out.zero_()
        h_raw = symm_mem.rendezvous(raw, group)
        h_out = symm_mem.rendezvous(out, group)
        self._raw2, self._out2 = raw, out
        self.raw, self.out = raw[0], out[0]
        self._symm = (h_raw, h_out, [int(p) for p in h_raw.buffer_ptrs], [int(p) for p in h_out.buffer_ptrs])
        self._raw_tick = 0
        self._out_tick = 0
        torch.cuda.synchronize(self.device)
        h_raw.barrier(channel=0)

    def set_scan(self, d_points: torch.Tensor) -> None:
        """The scan of the next step(s). One GPU with fused stages: only remembered here — step() prepares it in the same
        launch that builds the pose matrices (the tensor must stay unchanged until then)."""
        if self.world == 1 and hasattr(self.stages, "update_fused"):
            self._pending_scan = d_points
            return
        self._pending_scan = None
        self.stages.set_scan(d_points)

    def step(self, particles: torch.Tensor, n: int, tf, u0: float):
        """particles: float32[>= n, 7] on this rank's device, identical on every rank. Returns (resampled particles
        float32[n_out, 7] — a view into an internal buffer, valid until the next step —, mean pose float32[6] tensor,
        n_out, weight_sum). particles[:, 6] is overwritten with the normalised weights (cuda_eval_particles.h:556)."""
        W, r = self.world, self.rank
        self._reserve(n)
        chunk, first, count = shard(n, W, r)
        ocap = output_capacity(n, W)
        ochunk = ocap // W
        which_out = self._select_out()
        self._check_not_aliased(particles, n, ocap)
        if W == 1 and getattr(self, "_pending_scan", None) is not None:
            out = self.out[:ocap]
            self.stages.update_fused(self._pending_scan, particles, n, tf, u0, out, ocap, self.mean)
            n_out, wsum = self.stages.check()
            if n_out > ocap:
                raise RuntimeError(f"resampling emits {n_out} particles, capacity is {ocap}")
            return out[:n_out], self.mean[:6], n_out, wsum
        if self._symm is not None:
            h_raw, h_out, raw_ptrs, out_ptrs = self._symm
            which = self._select_raw()
            out = self.out[:ocap]
            raw_off = which * self._raw2.shape[1] * 4
            out_off = which_out * self._out2.shape[1] * 28 + r * ochunk * 28
            # The kernels themselves do the "all-gather" (P2P stores into every peer's copy); the two barriers only say
            # "all weights have landed everywhere" and "all resampled slices have landed everywhere".
            self.stages.eval_peers(particles, n, first, count, tf, self.raw, [p + raw_off for p in raw_ptrs])
            h_raw.barrier(channel=0)
            self.stages.normalize(particles, n, self.raw, self.mean)
            self.stages.draw_peers(particles, n, u0, r * ochunk, ochunk, out[r * ochunk:(r + 1) * ochunk], [p + out_off for p in out_ptrs])
            h_out.barrier(channel=0)
        else:
            out = self.out[:ocap]
            self.stages.eval(particles, n, first, count, tf, self.raw)
            if W > 1:
                self._all_gather(self.raw[:W * chunk], self.raw[r * chunk:(r + 1) * chunk])
            self.stages.normalize(particles, n, self.raw, self.mean)
            self.stages.draw(particles, n, u0, r * ochunk, ochunk, out[r * ochunk:(r + 1) * ochunk])
            if W > 1:
                self._all_gather(out.view(-1), out[r * ochunk:(r + 1) * ochunk].reshape(-1))
        n_out, wsum = self.stages.check()
        if n_out > ocap:
            raise RuntimeError(f"resampling emits {n_out} particles, capacity is {ocap}")
        return out[:n_out], self.mean[:6], n_out, wsum

    def _select_raw(self) -> int:
        """Alternate between the two symmetric-memory copies of the weight vector (every update that evaluates)."""
        which = self._raw_tick & 1
        self._raw_tick += 1
        self.raw = self._raw2[which]
        return which

    def _select_out(self) -> int:
        """Alternate between the two copies of the output particle buffer (every update that resamples): the set returned
        by step k lives in the copy step k + 1 does NOT write, so it may be passed back in as `particles`."""
        which = self._out_tick & 1
        self._out_tick += 1
        self.out = self._out2[which]
        return which

    def _check_not_aliased(self, particles: torch.Tensor, n: int, ocap: int) -> None:
        """The draw kernel reads parents from `particles` while it (and the exchange) writes `self.out`: overlapping the two
        would corrupt the resampled set silently."""
        if particles.device != self.out.device:
            return
        a0, a1 = particles.data_ptr(), particles.data_ptr() + n * 28
        b0, b1 = self.out.data_ptr(), self.out.data_ptr() + ocap * 28
        if a0 < b1 and b0 < a1:
            raise RuntimeError("particles overlaps the output buffer this step writes (pass a copy, or the set the previous step returned)")

    def evaluate_only(self, particles: torch.Tensor, n: int, tf):
        """Evaluation + normalisation without resampling (what the reference's evaluate() covers). Returns
        (mean pose tensor, weight_sum); normalised weights are left in particles[:, 6] on every rank."""
        W, r = self.world, self.rank
        self._reserve(n)
        chunk, first, count = shard(n, W, r)
        if getattr(self, "_pending_scan", None) is not None:      # the staged calls below need the prepared scan
            self.stages.set_scan(self._pending_scan)
            self._pending_scan = None
        if self._symm is not None:
            h_raw, _, raw_ptrs, _ = self._symm
            raw_off = self._select_raw() * self._raw2.shape[1] * 4
            self.stages.eval_peers(particles, n, first, count, tf, self.raw, [p + raw_off for p in raw_ptrs])
            h_raw.barrier(channel=0)
        else:
            self.stages.eval(particles, n, first, count, tf, self.raw)
            if W > 1:
                self._all_gather(self.raw[:W * chunk], self.raw[r * chunk:(r + 1) * chunk])
        self.stages.normalize(particles, n, self.raw, self.mean)
        _, wsum = self.stages.check()
        return self.mean[:6], wsum
