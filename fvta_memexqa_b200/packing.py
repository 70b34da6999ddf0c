"""Host-side packing of a feed dict for the packed reader (fvta_reader_fwd/bwd).

The reference feeds padded id tensors and lets TensorFlow run the LSTM over every padded
position (`model_v2.py:694-823`; ~75 % of the MemexQA positions are padding).  Lengths are
known on the HOST when the feed is built (`get_feed_dict`, model_v2.py:1099-1565), so the host
can hand the device (a) the valid tokens only, sequence after sequence, for all seven text call
sites at once, and (b) for every sequence the row of the output arena where its outputs belong
-- which IS its place in the context tensor ``hall[N,K,M,JMAX,2d]`` of model_v2.py:862-914.
Token counts are rounded up to a bucket so that CUDA graphs can be reused across batches.

`dynamic_rnn(sequence_length=len)` consumes the first ``len`` positions of a row, whatever the
mask pattern is; packing therefore takes the length-prefix of each row (identical to the mask
for the prefix masks the reference builds).
"""
import torch

TEXT_SITES = ("q", "at", "ad", "when", "where", "pts", "choices")


def _round_up(x, m):
    return (x + m - 1) // m * m


def build_pack(feed, config, bucket=512, min_text_cap=0, min_img_cap=0):
    """feed: placeholder-name -> CPU tensor.  Returns (tensors: dict name -> CPU tensor,
    meta: dict of python ints).  ``min_text_cap`` / ``min_img_cap`` pin the packed sizes (data
    parallel runs pin them to the global maxima so every rank replays the same single graph)."""
    N = feed["q"].shape[0]
    M = feed["at"].shape[1]
    JXA, JD = feed["at"].shape[2], feed["ad"].shape[2]
    JT, JG = feed["when"].shape[2], feed["where"].shape[2]
    JI, JXP = feed["pts"].shape[2], feed["pts"].shape[3]
    JQ = feed["q"].shape[1]
    C, JA = feed["choices"].shape[1], feed["choices"].shape[2]
    JMAX = max(JXA, JD, JT, JG, JI * JXP, JI)
    K = 5 if config.no_photo else 6
    T = M * JMAX
    hall_rows = N * K * T
    hq_row0 = hall_rows
    arena_rows = hall_rows + N * JQ

    n_idx = torch.arange(N).view(N, 1)
    m_idx = torch.arange(M).view(1, M)

    def hall_row(k):                       # first arena row of sequence (n, m) of modality k
        return ((n_idx * K + k) * M + m_idx) * JMAX

    seq_out = {
        "q": hq_row0 + torch.arange(N) * JQ,
        "at": hall_row(0).reshape(-1), "ad": hall_row(1).reshape(-1),
        "when": hall_row(2).reshape(-1), "where": hall_row(3).reshape(-1),
        "pts": (hall_row(4).unsqueeze(-1) + torch.arange(JI).view(1, 1, JI) * JXP).reshape(-1),
        "choices": torch.full((N * C,), -1, dtype=torch.int64),     # only the final state is consumed (:812)
    }
    ids, cids, lens_all, outs = [], [], [], []
    ranges = {}
    b0 = 0
    use_char = config.use_char
    for site in TEXT_SITES:
        mask = feed[site + "_mask"]
        J = mask.shape[-1]
        m2 = mask.reshape(-1, J)
        lens = m2.sum(1)
        prefix = torch.arange(J).view(1, J) < lens.view(-1, 1)
        ids.append(feed[site].reshape(-1, J)[prefix].long())
        if use_char:
            c = feed[site + "_c"]
            cids.append(c.reshape(-1, J, c.shape[-1])[prefix].to(torch.int32))
        lens_all.append(lens)
        outs.append(seq_out[site].long())
        ranges[site] = (b0, b0 + m2.shape[0])
        b0 += m2.shape[0]
    lens_t = torch.cat(lens_all)
    ntok = int(lens_t.sum())
    cap = max(bucket, _round_up(ntok, bucket), min_text_cap)
    tok_ids = torch.zeros(cap, dtype=torch.int64)
    tok_ids[:ntok] = torch.cat(ids)
    t = {"text_tok_ids": tok_ids,
         "text_seq_len": lens_t.to(torch.int32),
         "text_seq_off": (torch.cumsum(lens_t, 0) - lens_t).to(torch.int32),
         "text_seq_out": torch.cat(outs)}
    if use_char:
        W = cids[0].shape[-1]
        tc = torch.zeros(cap, W, dtype=torch.int32)
        tc[:ntok] = torch.cat(cids)
        t["text_tok_cids"] = tc
    # image reader: sequences (n, m) over photos; tokens are rows of image_emb_mat
    pm = feed["pis_mask"].reshape(-1, JI)
    plens = pm.sum(1)
    pprefix = torch.arange(JI).view(1, JI) < plens.view(-1, 1)
    pidx = feed["pis"].reshape(-1, JI)[pprefix].long()
    nimg = int(plens.sum())
    icap = max(64, _round_up(nimg, 64), min_img_cap)
    tp = torch.zeros(icap, dtype=torch.int64)
    tp[:nimg] = pidx
    t["img_tok_pidx"] = tp
    t["img_seq_len"] = plens.to(torch.int32)
    t["img_seq_off"] = (torch.cumsum(plens, 0) - plens).to(torch.int32)
    t["img_seq_out"] = (hall_row(5).reshape(-1) if K == 6 else torch.full((N * M,), -1)).long()
    # context-tensor mask [N,K,T] (model_v2.py:890-912), built here instead of pad+stack on device
    hm = torch.zeros(N, K, M, JMAX, dtype=torch.bool)
    hm[:, 0, :, :JXA] = feed["at_mask"]
    hm[:, 1, :, :JD] = feed["ad_mask"]
    hm[:, 2, :, :JT] = feed["when_mask"]
    hm[:, 3, :, :JG] = feed["where_mask"]
    hm[:, 4, :, :JI * JXP] = feed["pts_mask"].reshape(N, M, JI * JXP)
    if K == 6:
        hm[:, 5, :, :JI] = feed["pis_mask"]
    t["hall_mask"] = hm.reshape(N, K, T)
    meta = dict(N=N, M=M, K=K, T=T, JMAX=JMAX, JQ=JQ, C=C, JXP=JXP, arena_rows=arena_rows, hall_rows=hall_rows,
                hq_row0=hq_row0, text_cap=cap, img_cap=icap, text_ntok=ntok, img_ntok=nimg,
                text_B=int(lens_t.numel()), img_B=int(plens.numel()),
                text_Jmax=max(JQ, JXA, JD, JT, JG, JXP, JA), img_Jmax=JI,
                q_range=ranges["q"], choices_range=ranges["choices"])
    return t, meta


def emb_touch_flags(feed, touch_rows, touch_len):
    """1.0 for every row of the sparse embedding tables (params.SPARSE_TABLES) that some id of the FED
    arrays looks up -- padding positions included: the reference feeds the padded id arrays to
    tf.nn.embedding_lookup (model_v2.py:534-546, :599-606), so NULL (id 0) is looked up whenever a batch
    has any padding, and TF's sparse Adadelta decays that row's accumulators even with a zero
    gradient.  Ids >= VW address the fed GloVe matrix (model_v2.py:596, :1335), which is not trained."""
    flags = torch.zeros(touch_len, dtype=torch.float32)
    for name, (base, rows) in touch_rows.items():
        suffix = "_c" if name.endswith("char_emb") else ""
        for site in TEXT_SITES:
            ids = feed[site + suffix].reshape(-1).long()
            flags[base + ids[ids < rows]] = 1.0
    return flags


# tensors the device step consumes besides the packed reader inputs
DEVICE_KEYS = ("q_mask", "y", "existing_emb_mat", "image_emb_mat")


def host_step_input(feed, config, min_text_cap=0, min_img_cap=0, touch_rows=None, touch_len=0, pin=True):
    """Everything the device step reads, as ONE host buffer (to_blob): the packed reader inputs, the
    question mask, the labels, whichever embedding tables are fed and -- in training -- the flags of the
    embedding rows this batch looks up.  Pure host work with no model state: an input-pipeline worker
    process runs it (pipeline.BatchLoader); Model.host_step_input is this with the model's settings."""
    host = {k: (torch.as_tensor(v) if not isinstance(v, torch.Tensor) and hasattr(v, "shape") else v)
            for k, v in feed.items()}
    tensors, meta = build_pack(host, config, min_text_cap=min_text_cap, min_img_cap=min_img_cap)
    for k in DEVICE_KEYS:
        if k in host and host[k] is not None:
            tensors[k] = host[k]
    if host.get("is_train", False) and touch_len:
        tensors["emb_touch"] = emb_touch_flags(host, touch_rows, touch_len)
    blob, index = to_blob(tensors, pin=pin)
    meta["blob_bytes"] = blob.numel()
    return {"blob": blob, "index": index, "meta": meta, "host": host}


def to_blob(tensors, pin=True):
    """Concatenate CPU tensors into ONE (pinned) byte buffer so a step's inputs travel in a single
    host->device copy.  Returns (blob uint8 [nbytes], index: name -> (offset, shape, dtype))."""
    index, off = {}, 0
    for k, v in tensors.items():
        v = v.contiguous()
        index[k] = (off, tuple(v.shape), v.dtype)
        off += (v.numel() * v.element_size() + 255) // 256 * 256
    blob = torch.empty(max(off, 256), dtype=torch.uint8)
    if pin and torch.cuda.is_available():
        blob = blob.pin_memory()
    for k, v in tensors.items():
        o, shape, dtype = index[k]
        n = v.numel() * v.element_size()
        blob[o:o + n] = v.contiguous().view(-1).view(torch.uint8)
    return blob, index


def blob_views(dev_blob, index):
    """Typed views into a device copy of the blob."""
    out = {}
    for k, (o, shape, dtype) in index.items():
        n = 1
        for s_ in shape:
            n *= s_
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        out[k] = dev_blob[o:o + nbytes].view(dtype).view(shape)
    return out
