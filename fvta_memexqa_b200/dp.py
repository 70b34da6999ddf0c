"""Data parallelism over the QA batch (north_star d): one process per GPU, every parameter
replicated, rank r trains on its own shard of the batch, and the ONE flat fp32 gradient
buffer (ParamStore.grad_flat) is sum-all-reduced by NCCL over NVLink once per step; the
1/world_size factor that turns the sum of per-rank batch-means into the global batch-mean
(model_v2.py:1090 reduce_mean) is folded into the Adadelta kernel's ``grad_scale``.

The reference is single-device (model_v2.py:14); this is the only place the path shards
(SURVEY section 8e).  Padded shapes must be pinned to the global batch maxima so that a
G-GPU step equals the 1-GPU step on the concatenated batch.
"""
import os

import torch
import torch.distributed as dist


class DataParallel:
    def __init__(self, backend=None, device=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = device
        self.p2p = None
        if self.world_size > 1 and not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                self.device = torch.device("cuda", self.local_rank)
                dist.init_process_group(backend, device_id=self.device)
            else:
                dist.init_process_group(backend)

    # ------------------------------------------------------------------ fused all-reduce + Adadelta
    def enable_p2p(self, ps):
        """Peer-mapped staging buffers for `fvta_p2p_allreduce_adadelta` (csrc/p2p_adadelta.cu): the
        gradient sum over the ranks and the optimizer update run as ONE kernel per rank over NVLink,
        replacing ncclAllReduce + fvta_adadelta.  torch's symmetric memory supplies the allocation and
        the exchange of the peer mappings (plumbing only); the data path is this repo's kernel.
        Returns False (and leaves the NCCL path in place) when peer mappings are unavailable."""
        self.p2p = None
        if self.world_size <= 1 or not torch.cuda.is_available():
            return False
        try:
            import torch.distributed._symmetric_memory as symm_mem
            total = ps.grad_buf.numel()
            stage = symm_mem.empty(2 * total, dtype=torch.float32, device=self.device)
            flags = symm_mem.empty(64, dtype=torch.int32, device=self.device)
            stage.zero_()
            flags.zero_()
            torch.cuda.synchronize()
            hs = symm_mem.rendezvous(stage, dist.group.WORLD)
            hf = symm_mem.rendezvous(flags, dist.group.WORLD)
            state = torch.zeros(4, dtype=torch.int32, device=self.device)
            self.p2p = dict(stage=stage, flags=flags, hs=hs, hf=hf, state=state,
                            stage_ptrs=int(hs.buffer_ptrs_dev), flag_ptrs=int(hf.buffer_ptrs_dev),
                            n=ps.flat.numel(), n_tail=ps.touch_len)
            dist.barrier()
            torch.cuda.synchronize()
            return True
        except Exception as e:      # no peer access / no symmetric-memory support on this system
            if self.rank == 0:
                print("dp: peer-memory all-reduce unavailable (%s: %s); using ncclAllReduce + fvta_adadelta"
                      % (type(e).__name__, e), flush=True)
            self.p2p = None
            return False

    def allreduce_adadelta(self, ps, lr, rho, eps):
        """One launch: sum of ps.grad_buf over the ranks (rank order) / world -> Adadelta on ps.flat."""
        import ctypes
        from . import _lib
        lib = _lib.load()
        p = self.p2p
        vp = ctypes.c_void_p
        rc = lib.fvta_p2p_allreduce_adadelta(
            self.world_size, self.rank, p["n"], p["n_tail"], vp(ps.grad_buf.data_ptr()), vp(p["stage_ptrs"]),
            vp(p["flag_ptrs"]), vp(p["state"].data_ptr()), vp(ps.flat.data_ptr()), vp(ps.accum.data_ptr()),
            vp(ps.accum_update.data_ptr()), vp(ps.update_mask.data_ptr()),
            vp(ps.touch_index.data_ptr()) if ps.touch_index is not None else None, lr, rho, eps,
            vp(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "fvta_p2p_allreduce_adadelta")

    def shard(self, n_global):
        """Half-open range of global sample indices this rank owns."""
        per = n_global // self.world_size
        rem = n_global % self.world_size
        lo = self.rank * per + min(self.rank, rem)
        return lo, lo + per + (1 if self.rank < rem else 0)

    def all_reduce_grads(self, grad_flat):
        """Sum-all-reduce the flat gradient buffer in place; returns the scale (1/world) the
        optimizer must apply."""
        if self.world_size > 1:
            dist.all_reduce(grad_flat, op=dist.ReduceOp.SUM)
        return 1.0 / self.world_size

    def mean_loss(self, loss):
        """The loss a rank reports is its LOCAL batch mean (no collective on the step's critical path for a
        number that is only logged); `global_mean` all-reduces it when a caller wants the global one."""
        return loss

    def global_mean(self, t):
        """All-reduced mean of a scalar tensor (reporting only)."""
        if self.world_size > 1:
            t = t.detach().clone()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t = t / self.world_size
        return t

    def pin_caps(self, model, batches):
        """Pin the packed-token sizes to the maximum over every batch of every rank, so that all
        ranks build ONE step graph and capture it in lockstep (a rank that captured an extra
        bucket would issue all-reduces its peers never join)."""
        caps = [0, 0]
        for b in batches:
            meta = model.get_feed_dict(b, is_train=True)["meta"]
            caps = [max(caps[0], meta["text_cap"]), max(caps[1], meta["img_cap"])]
        if self.world_size > 1:
            n = int(model.config.batch_size)
            t = torch.tensor(caps + [n, -n], dtype=torch.int64, device=self.device if self.device is not None else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            caps = [int(v) for v in t.tolist()[:2]]
            # the gradient average is sum / world: right only when every rank's loss is a mean over the same N
            # (the static config.batch_size -- padded rows count in the mean, model_v2.py:1103)
            if int(t[2]) != -int(t[3]):
                raise ValueError("data parallel: every rank must use the same config.batch_size")
        model.pin_text_cap, model.pin_img_cap = caps
        return caps

    def broadcast_params(self, flat):
        if self.world_size > 1:
            dist.broadcast(flat, src=0)

    def barrier(self):
        if self.world_size > 1:
            dist.barrier()

    def max_over_ranks(self, value):
        if self.world_size == 1:
            return value
        dev = self.device if self.device is not None else "cpu"
        t = torch.tensor([value], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
