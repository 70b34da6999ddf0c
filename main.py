#!/usr/bin/env python
# coding=utf-8
"""Train / test driver with the reference's command line.

Mirrors the interface of the reference's ``main.py``: ``python main.py prepropath outbasepath
[flags]`` with every flag of main.py:28-147, ``train(config)`` (main.py:278-458: periodic
save + validation, best-checkpoint tracking, train_sum.txt), ``test(config)``
(main.py:462-637: answers.json, optional yps.json / visualisation pickles / weights dump) and
``read_data`` (main.py:215-274: ``{train,val,test}_{data,shared}.p`` pickles and the run-dir
``shared.p`` vocabulary side-car).  The model, step runners and kernels come from
``fvta_memexqa_b200``; TensorFlow sessions/savers are replaced by torch checkpoints of the flat
parameter buffer (``save-<step>.pt`` / ``save-best-<step>.pt``).

Added: ``--synthetic N`` builds a MemexQA-shaped dataset in the reference's on-disk format in
memory (the real dataset cannot be downloaded here), ``--num_choice``, ``--seed``.
"""
import json
import math
import os
import pickle
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from fvta_memexqa_b200.config import get_args  # noqa: E402
from fvta_memexqa_b200.data import (Dataset, update_config, getAnswers, getAnswers_yp, getEvalScore,  # noqa: E402
                                    make_synthetic_raw, add_vocab)


def sec2time(secs):
    """utils.py:20-29."""
    m, s = divmod(secs, 60)
    h, m = divmod(m, 60)
    return ("%02d:%02d:%.3f" if s >= 10.0 else "%02d:%02d:0%.3f") % (h, m, s)


def read_data(config, datatype, loadExistModelShared=False):
    """main.py:215-274."""
    if config.synthetic > 0:
        n = config.synthetic if datatype == "train" else max(config.batch_size, config.synthetic // 4)
        data, shared = make_synthetic_raw(n, seed=config.seed + {"train": 0, "val": 1, "test": 2}[datatype],
                                          feat_dim=config.image_feat_dim)
    else:
        with open(os.path.join(config.prepropath, "%s_data.p" % datatype), "rb") as f:
            data = pickle.load(f, encoding="latin1")       # the reference writes Python-2 pickles
        with open(os.path.join(config.prepropath, "%s_shared.p" % datatype), "rb") as f:
            shared = pickle.load(f, encoding="latin1")
    num_examples = len(data["q"])
    valid_idxs = range(num_examples)
    print("loaded %s/%s data points for %s" % (len(valid_idxs), num_examples, datatype))
    model_shared_path = os.path.join(config.outpath, "shared.p")
    if loadExistModelShared:
        with open(model_shared_path, "rb") as f:
            add_vocab(shared, config, pickle.load(f))
    else:
        side = add_vocab(shared, config)
        with open(model_shared_path, "wb") as f:
            pickle.dump(side, f)
    feat = next(iter(shared["pid2feat"].values()))
    assert config.image_feat_dim == feat.shape[0], "image dim is not %s, it is %s" % (config.image_feat_dim, feat.shape[0])
    return Dataset(data, datatype, shared=shared, valid_idxs=valid_idxs)


def initialize(load, load_best, model, config):
    """main.py:640-665: fresh initialisation (done by the Model) or restore the latest / best
    checkpoint of the run directory."""
    import torch
    if not load:
        return
    d = config.save_dir_best if load_best else config.save_dir
    ckpts = sorted((f for f in os.listdir(d) if f.endswith(".pt")),
                   key=lambda f: int(f.rsplit("-", 1)[1][:-3])) if os.path.isdir(d) else []
    if not ckpts:
        raise Exception("Model not exists")
    path = os.path.join(d, ckpts[-1])
    print("restoring model from %s" % path)
    sd = torch.load(path, map_location="cpu")
    model.params.load_state_dict(sd["params"])
    model.global_step = sd["global_step"]


def _save(model, path_prefix, step, keep=5):
    import torch
    path = "%s-%d.pt" % (path_prefix, step)
    torch.save({"params": model.params.state_dict(), "global_step": model.global_step}, path)
    d, base = os.path.split(path_prefix)
    old = sorted((f for f in os.listdir(d) if f.startswith(base + "-") and f.endswith(".pt")),
                 key=lambda f: int(f.rsplit("-", 1)[1][:-3]))
    for f in old[:-keep]:                                  # tf.train.Saver(max_to_keep=5), main.py:314-315
        os.remove(os.path.join(d, f))


def cal_total_param(model):
    """main.py:670-678."""
    return sum(int(np.prod(shape)) for (_, _, shape) in model.params.offsets.values())


def _evaluate(tester, dataset, config, num_batches):
    pred, gt = {}, {}
    for evalbatch in dataset.get_batches(config.batch_size, num_steps=num_batches, shuffle=False, cap=True):
        yp = tester.step(None, evalbatch)
        p, g = getAnswers(yp, evalbatch)
        pred.update(p)
        gt.update(g)
    return getEvalScore(pred, gt)


def train(config):
    """main.py:278-458."""
    from fvta_memexqa_b200.model import select_get_model
    from fvta_memexqa_b200.trainer import Trainer
    from fvta_memexqa_b200.tester import Tester
    self_summary_strs = []
    train_data = read_data(config, "train", config.load)
    val_data = read_data(config, "val", True)
    update_config(config, [train_data, val_data], showMeta=True)
    word2vec_dict, word2idx_dict = train_data.shared["word2vec"], train_data.shared["word2idx"]
    idx2vec_dict = {word2idx_dict[w]: vec for w, vec in word2vec_dict.items() if w in word2idx_dict}
    rng = np.random.RandomState(config.seed)
    config.emb_mat = np.array([idx2vec_dict[i] if i in idx2vec_dict else rng.randn(config.word_emb_size)
                               for i in range(config.word_vocab_size)], dtype="float32")   # main.py:308
    model = select_get_model(config)(config)
    trainer = Trainer(model, config)
    tester = Tester(model, config)
    s = "total parameters: %s" % cal_total_param(model)
    print(s)
    self_summary_strs.append(s)
    initialize(config.load, config.load_best, model, config)
    last_time = time.time()
    num_steps = int(math.ceil(train_data.num_examples / float(config.batch_size))) * config.num_epochs
    s = " batch_size:%s, epoch:%s,total step:%s,eval/save every %s steps" % (
        config.batch_size, config.num_epochs, num_steps, config.save_period)
    print(s)
    self_summary_strs.append(s)
    best = {"acc": 0.0, "step": -1}
    finalAcc = None
    pyrng = random.Random(config.seed)
    for batch in train_data.get_batches(config.batch_size, num_steps=num_steps, rng=pyrng):
        global_step = model.global_step + 1
        loss, _, _ = trainer.step(None, batch)
        if global_step % config.save_period == 0:
            duration = time.time() - last_time
            sec_per_step = duration / config.save_period
            last_time = time.time()
            _save(model, config.save_dir_model, global_step)
            evalAcc = _evaluate(tester, val_data, config, config.val_num_batches)
            print("step:%s/%s (epoch:%.3f), took %s, loss:%s, val acc:%s, best:%s at step %s (%.3f secs/step, eta:%s)" % (
                global_step, num_steps, config.num_epochs * global_step / float(num_steps), sec2time(duration), loss,
                evalAcc, best["acc"], best["step"], sec_per_step, sec2time((num_steps - global_step) * sec_per_step)))
            if evalAcc > best["acc"]:
                best.update(acc=evalAcc, step=global_step)
                _save(model, config.save_dir_best_model, global_step)
            finalAcc = evalAcc
    if model.global_step % config.save_period != 0:
        _save(model, config.save_dir_model, model.global_step)
    s = "best eval on val Accuracy: %s, at step %s, final step %s, acc is %s" % (
        best["acc"], best["step"], model.global_step, finalAcc)
    print(s)
    self_summary_strs.append(s)
    if config.write_self_sum:
        with open(config.self_summary_path, "w") as f:
            f.writelines("%s\n" % x for x in self_summary_strs)
    return model


def test(config):
    """main.py:462-637."""
    from fvta_memexqa_b200.model import select_get_model
    from fvta_memexqa_b200.tester import Tester
    if config.is_test_on_val:
        test_data = read_data(config, "val", True)
        print("total val samples:%s" % test_data.num_examples)
    else:
        test_data = read_data(config, "test", True)
        print("total test samples:%s" % test_data.num_examples)
    update_config(config, [test_data], showMeta=True)
    config.emb_mat = None
    model = select_get_model(config)(config)
    initialize(True, config.load_best, model, config)
    if config.is_save_weights:
        np.savez(os.path.join(config.weights_path, "weights.npz"),
                 **{k.replace("/", "."): v.detach().cpu().numpy() for k, v in model.params.as_dict().items()})
        with open(os.path.join(config.weights_path, "all.txt"), "w") as f:
            f.writelines("%s\n" % k for k in model.params.offsets)
    tester = Tester(model, config)
    num_steps = int(math.ceil(test_data.num_examples / float(config.batch_size)))
    id2predanswers, id2realanswers, id2yp = {}, {}, {}
    for count, batch in enumerate(test_data.get_batches(config.batch_size, num_steps=num_steps, shuffle=False), 1):
        if config.is_save_vis:
            out = tester.step_vis(None, batch)
            yp = out[0]
            with open(os.path.join(config.vis_path, "%s.p" % count), "wb") as f:
                pickle.dump({"yp": yp, "C": out[1], "C_win": out[2], "att_logits": out[3], "q_att_logits": out[4],
                             "data": batch[1].data}, f)
        else:
            yp = tester.step(None, batch)
        if config.get_yp:
            pred, gt, yps = getAnswers_yp(yp, batch)
            id2yp.update(yps)
        else:
            pred, gt = getAnswers(yp, batch)
        id2predanswers.update(pred)
        id2realanswers.update(gt)
    acc = getEvalScore(id2predanswers, id2realanswers)
    print("done, got %s answers, accuracy:%s" % (len(id2predanswers), acc))
    with open(os.path.join(config.val_path, "answers.json"), "w") as f:
        json.dump(id2predanswers, f)
    if config.get_yp:
        with open(os.path.join(config.val_path, "yps.json"), "w") as f:
            json.dump({k: [float(x) for x in v] for k, v in id2yp.items()}, f)
    return acc


if __name__ == "__main__":
    config = get_args()
    # main.py:684-707: the key lists printed by train()/update_config
    config.thresmeta = ["word_count_thres", "char_count_thres", "sent_album_title_size_thres",
                        "sent_photo_title_size_thres", "sent_des_size_thres", "sent_when_size_thres",
                        "sent_where_size_thres", "answer_size_thres", "question_size_thres", "num_photos_thres",
                        "num_albums_thres", "word_size_thres"]
    config.maxmeta = ["max_num_albums", "max_num_photos", "max_sent_album_title_size", "max_sent_photo_title_size",
                      "max_sent_des_size", "max_when_size", "max_where_size", "max_answer_size",
                      "max_question_size", "max_word_size"]
    if config.is_pack_model:
        raise NotImplementedError("--is_pack_model re-saves a TensorFlow graph with named collections "
                                  "(main.py:498-576); there is no TF graph here")
    if config.is_train:
        train(config)
    else:
        test(config)
