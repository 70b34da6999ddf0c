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
        ps.grad_flat.zero_()
        loss, _, _ = m.forward(dfeed, is_train=True)
        loss.backward()
        scale = 1.0
        if self.dp is not None:
            scale = self.dp.all_reduce_grads(ps.grad_flat)
            loss = self.dp.mean_loss(loss)
        ops.adadelta_step(ps.flat, ps.grad_flat, ps.accum, ps.accum_update, ps.update_mask,
                          self.lr, self.rho, self.eps, grad_scale=scale)
        m.global_step += 1
        self.global_step = m.global_step
        return loss

    def step(self, sess, batch, get_summary=False):
        """trainer.py:30-40.  ``sess`` is accepted for signature parity (the reference asserts
        a tf.Session; here it is an opaque context and may be None)."""
        batchIdx, batch_data = batch
        feed = self.model.get_feed_dict(batch_data, is_train=True)
        dfeed = self.model.to_device(feed)
        loss = self._device_step(dfeed)
        self.loss = loss
        return float(loss.detach()), None, None   # (loss, summary, train_op)
