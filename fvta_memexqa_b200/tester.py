"""Inference step runner with the reference's interface.

Mirrors the interface of the reference's ``tester.py:8-43``: ``Tester(model, config, sess=None)``,
``step(sess, batch) -> yp[:num_examples]`` (numpy [num_examples, num_choice]) and
``step_vis(sess, batch)`` returning the 28 arrays the reference fetches, the same eight of
them trimmed to the real examples of a short last batch.
"""
import numpy as np
import torch


class Tester:
    def __init__(self, model, config, sess=None):
        self.config = config
        self.model = model
        self.yp = None

    def step(self, sess, batch):
        """tester.py:16-25."""
        batchIdxs, batch_data = batch
        feed = self.model.get_feed_dict(batch_data, is_train=False)
        dfeed = self.model.to_device(feed)
        with torch.no_grad():
            _, _, yp = self.model.forward(dfeed, is_train=False)
        self.yp = yp
        return yp[:batch_data.num_examples].cpu().numpy()

    def trim(self, input_s, num):
        """tester.py:27-28."""
        return [one[:num] if isinstance(one, np.ndarray) and one.ndim > 0 else -1 for one in input_s]

    def step_vis(self, sess, batch):
        """tester.py:30-43: everything needed for visualisation."""
        batchIdxs, batch_data = batch
        m = self.model
        feed = m.get_feed_dict(batch_data, is_train=False)
        dfeed = m.to_device(feed)
        with torch.no_grad():
            _, _, yp = m.forward(dfeed, is_train=False, want_vis=True)

        def np_(x):
            if isinstance(x, torch.Tensor):
                return x.detach().cpu().numpy()
            return x

        n = batch_data.num_examples
        yp = np_(yp)[:n]
        C, C_win, att_logits, q_att_logits = np_(m.C), m.C_win, np_(m.att_logits), np_(m.q_att_logits)
        host = dfeed["_host"]
        masks = {k: np_(host[k]) for k in ("at_mask", "ad_mask", "when_mask", "where_mask", "pts_mask",
                                           "pis_mask", "q_mask")}
        C, C_win, att_logits, q_att_logits, at_mask, pts_mask, pis_mask, q_mask = self.trim(
            [C, C_win, att_logits, q_att_logits, masks["at_mask"], masks["pts_mask"], masks["pis_mask"],
             masks["q_mask"]], n)
        ids = {k: np_(host[k]) for k in ("at", "ad", "when", "where", "pts", "pis", "q")}
        return (yp, C, C_win, att_logits, q_att_logits, at_mask, masks["ad_mask"], masks["when_mask"],
                masks["where_mask"], pts_mask, pis_mask, q_mask, m.hat_len, m.had_len, m.hwhen_len,
                m.hwhere_len, m.hpts_len, m.hpis_len, m.JXP, np_(m.warp_h), np_(m.hall), ids["at"],
                ids["ad"], ids["when"], ids["where"], ids["pts"], ids["pis"], ids["q"])
