"""Training step runner with the reference's interface.

Mirrors the interface of the reference's ``trainer.py:9-40``: ``Trainer(model, config)`` and
``step(sess, batch, get_summary=False) -> (loss, summary, train_op)`` with
``batch = (idxs, Dataset)``.  One step = host feed -> device, forward, backward, (data-parallel
all-reduce of the flat gradient buffer), Adadelta over the flat parameter buffer, global_step++.

``tf.train.AdadeltaOptimizer(config.init_lr)`` [TF defaults rho=0.95, eps=1e-8] becomes one
launch of ``fvta_adadelta`` over ``ParamStore.flat``.
"""
import torch

from . import ops


class StepHandle:
    """The loss of an enqueued step: its device->host copy was issued right behind the step; ``result()``
    waits for that copy only (not for whatever was enqueued after it)."""

    def __init__(self, host, slot, event):
        self._host, self._slot, self._event = host, slot, event

    def result(self):
        self._event.synchronize()
        return float(self._host[self._slot])


class Trainer:
    def __init__(self, model, config, dp=None):
        self.config = config
        self.model = model
        self.global_step = model.global_step
        self.lr = config.init_lr
        self.rho, self.eps = 0.95, 1e-8
        self.loss = None
        self.summary = model.summary
        self.dp = dp            # fvta_memexqa_b200.dp.DataParallel or None

    def _device_step(self, dfeed):
        """forward + backward + (all-reduce) + Adadelta on device tensors; returns the loss tensor."""
        m = self.model
        ps = m.params
        ps.grad_buf.zero_()
        if ps.touch is not None:
            # which embedding rows this batch looks up (tail of the gradient buffer: all-reduced with it)
            ps.touch.copy_(dfeed["emb_touch"])
        loss, _, _ = m.forward(dfeed, is_train=True)
        loss.backward()
        # the image side of the model runs on a second stream, forward and backward (Model.forward); its
        # gradients land in the flat buffer through sinks, which autograd does not wait for: join here
        torch.cuda.current_stream().wait_stream(ops._reader_side_stream(m.device))
        loss = loss.detach()
        m.release_graph()
        scale = 1.0
        if self.dp is not None and getattr(self.dp, "p2p", None) is not None:
            # all-reduce over NVLink peer memory fused with the optimizer update: one launch, no NCCL
            self.dp.allreduce_adadelta(ps, self.lr, self.rho, self.eps)
            m.global_step += 1
            self.global_step = m.global_step
            return self.dp.mean_loss(loss)
        if self.dp is not None:
            scale = self.dp.all_reduce_grads(ps.grad_buf)
            loss = self.dp.mean_loss(loss)
        ops.adadelta_step(ps.flat, ps.grad_flat, ps.accum, ps.accum_update, ps.update_mask,
                          self.lr, self.rho, self.eps, grad_scale=scale, touch_index=ps.touch_index,
                          touch=ps.touch)
        m.global_step += 1
        self.global_step = m.global_step
        return loss

    # ---------------------------------------------------------------- CUDA-graph execution
    def enable_graph(self, example_feed=None, warmup=3):
        """Run the whole device step (forward, backward, all-reduce, Adadelta) as ONE CUDA graph
        over static input buffers.  At the reference shape the step is ~100 short kernels and
        launch-latency bound, so replaying a graph is the difference between the GPU and the
        Python interpreter setting the pace.  Requires pinned padded shapes; graphs are keyed by
        the packed-token buckets (packing.build_pack) and captured on first use."""
        from collections import OrderedDict
        self._graphs = OrderedDict()
        self._graph_warmup = warmup
        self.max_graphs = 8          # bounded: each entry retains a static input blob and a memory pool
        if example_feed is not None:
            self._capture(self.model.to_device(example_feed))
        return self

    @staticmethod
    def _graph_key(meta, index):
        """Everything a captured step bakes in as a constant: every shape-bearing field of the packing
        meta (N, M, K, T, JMAX, JQ, arena rows, sequence counts, ranges, packed capacities, ...) and the
        layout of the step input (name, offset, shape, dtype of every tensor in the blob -- this is
        where e.g. the photo-table row count lives).  Only the true token counts may differ: the
        kernels read the per-sequence lengths from the blob."""
        m = tuple(sorted((k, v) for k, v in meta.items() if k not in ("text_ntok", "img_ntok")))
        lay = tuple(sorted((k, int(o), tuple(shape), str(dtype)) for k, (o, shape, dtype) in index.items()))
        return (m, lay)

    def _resident_key(self):
        """Device-resident tables (Model.set_resident_tables) are baked into a captured step by address."""
        return tuple(sorted((k, v.data_ptr(), tuple(v.shape)) for k, v in getattr(self.model, "resident", {}).items()))

    def _capture(self, dfeed):
        from . import _lib
        lib = _lib.load()
        m = self.model
        ps = m.params
        meta = dfeed["_meta"]
        from .packing import blob_views
        hsi_index = {k: (v.data_ptr() - dfeed["_blob"].data_ptr(), tuple(v.shape), v.dtype)
                     for k, v in dfeed.items() if isinstance(v, torch.Tensor) and k != "_blob"}
        key = self._graph_key(meta, hsi_index) + (self._resident_key(),)
        while len(self._graphs) >= self.max_graphs:
            self._graphs.popitem(last=False)
        static_blob = dfeed["_blob"].clone()
        static = blob_views(static_blob, hsi_index)
        static.update({k: v for k, v in dfeed.items() if not isinstance(v, torch.Tensor)})
        static["_blob"] = static_blob
        # warm-up and capture must not move the weights
        snap = [t.clone() for t in (ps.flat, ps.accum, ps.accum_update)]
        step0 = m.global_step
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self._graph_warmup):
                self._device_step(static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = lib.fvta_launch_count()
        with torch.cuda.graph(graph):
            loss = self._device_step(static)
        self.launches_per_step = int(lib.fvta_launch_count() - n0)   # library kernels in the graph
        for t, s0 in zip((ps.flat, ps.accum, ps.accum_update), snap):
            t.copy_(s0)
        m.global_step = step0
        self.global_step = step0
        self._graphs[key] = (graph, static, loss)
        return self._graphs[key]

    # ---------------------------------------------------------------- input prefetch
    def prefetch(self, batch):
        """Start the host->device copy of the NEXT batch's step input on a copy stream, so that it runs
        under the current step's kernels instead of in front of the next step's (5 MB over PCIe is
        ~100 us, a sixth of the C2 step).  `batch` = (idxs, Dataset) like `step`.  The following
        `step(batch)` picks the staged copy up with one device-to-device copy into the graph's static
        input.  Optional: a step that was not prefetched copies from the host as before."""
        feed = self.model.get_feed_dict(batch[1], is_train=True)
        if "blob" not in feed:
            return
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
            self._stage = [None, None]         # two staging buffers: one being filled, one being consumed
            self._stage_free = [None, None]    # event: the compute stream has consumed the slot
            self._stage_slot = 0
            self._staged = {}                  # id(host blob) -> (host blob, slot, ready event)
        slot = self._stage_slot
        self._stage_slot ^= 1
        src = feed["blob"]
        if self._stage[slot] is None or self._stage[slot].numel() != src.numel():
            self._stage[slot] = torch.empty(src.numel(), dtype=torch.uint8, device=self.model.device)
            self._stage_free[slot] = None
        with torch.cuda.stream(self._copy_stream):
            if self._stage_free[slot] is not None:
                self._copy_stream.wait_event(self._stage_free[slot])
            self._stage[slot].copy_(src, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self._copy_stream)
        for k in [k for k, v in self._staged.items() if v[1] == slot]:
            del self._staged[k]                # the slot's previous occupant was never consumed
        self._staged[id(src)] = (src, slot, ready)

    def _staged_input(self, src):
        """Device copy of host blob `src` if `prefetch` staged exactly this blob, else None."""
        st = getattr(self, "_staged", {}).pop(id(src), None)
        if st is None or st[0] is not src:
            return None
        _, slot, ready = st
        torch.cuda.current_stream().wait_event(ready)
        return slot

    def graph_step(self, feed):
        """Replay the captured step on a new batch (host feed or device step input).  Returns the
        loss tensor (device)."""
        dfeed = None
        if "blob" in feed:
            key = feed.get("_graph_key")
            if key is None:          # cached on the host step input: computed once per batch object
                key = feed["_graph_key"] = self._graph_key(feed["meta"], feed["index"])
            key = key + (self._resident_key(),)
        else:
            dfeed = self.model.to_device(feed)
            index = {k: (v.data_ptr() - dfeed["_blob"].data_ptr(), tuple(v.shape), v.dtype)
                     for k, v in dfeed.items() if isinstance(v, torch.Tensor) and k != "_blob"}
            key = self._graph_key(dfeed["_meta"], index) + (self._resident_key(),)
        entry = self._graphs.get(key)
        if entry is None:
            dfeed = dfeed if dfeed is not None else self.model.to_device(feed)
            entry = self._capture(dfeed)
        else:
            self._graphs.move_to_end(key)
        graph, static, loss = entry
        src = feed["blob"] if "blob" in feed else dfeed["_blob"]     # pinned host blob: copy it straight in
        if src.numel() != static["_blob"].numel():
            raise ValueError("graph step: step input is %d bytes, captured %d (pin the padded shapes)"
                             % (src.numel(), static["_blob"].numel()))
        slot = self._staged_input(src) if "blob" in feed else None
        if slot is not None:         # prefetched: device-to-device, the PCIe copy already happened
            static["_blob"].copy_(self._stage[slot], non_blocking=True)
            done = torch.cuda.Event()
            done.record()
            self._stage_free[slot] = done
        else:
            static["_blob"].copy_(src, non_blocking=True)
        graph.replay()
        self.model.global_step += 1
        self.global_step = self.model.global_step
        return loss

    def step_async(self, sess, batch):
        """Enqueue one training step and return its loss as a DEVICE tensor without waiting for it: the
        caller fetches and prefetches the next batch while the step runs, then reads the loss
        (``float(loss)``).  ``step`` is this plus the read-back."""
        batchIdx, batch_data = batch
        feed = self.model.get_feed_dict(batch_data, is_train=True)
        if getattr(self, "_graphs", None) is not None:
            loss = self.graph_step(feed)
        else:
            loss = self._device_step(self.model.to_device(feed))
        self.loss = loss
        return loss

    def step_handle(self, sess, batch):
        """``step_async`` plus an asynchronous read-back of the loss into pinned host memory: the training
        loop can enqueue step i+1 before it looks at the loss of step i, so the device never waits for the
        interpreter between steps.  Returns a StepHandle; up to four may be outstanding."""
        loss = self.step_async(sess, batch)
        if getattr(self, "_loss_host", None) is None:
            self._loss_host = torch.zeros(4, dtype=torch.float32).pin_memory()
            self._loss_slot = 0
        k = self._loss_slot
        self._loss_slot = (k + 1) % 4
        self._loss_host[k:k + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return StepHandle(self._loss_host, k, ev)

    def step(self, sess, batch, get_summary=False):
        """trainer.py:30-40.  ``sess`` is accepted for signature parity (the reference asserts
        a tf.Session; here it is an opaque context and may be None)."""
        loss = self.step_async(sess, batch)
        return float(loss.detach()), None, None   # (loss, summary, train_op)
