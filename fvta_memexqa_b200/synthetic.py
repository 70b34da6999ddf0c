"""Synthetic MemexQA-shaped inputs (there is no dataset offline).

Two generators, both seeded and CPU-side (the caller moves tensors to the device):

* :func:`make_feed` -- a full model feed: one tensor per placeholder of the reference's
  ``Model`` (model_v2.py:420-477: at, at_c, at_mask, ..., pis, pis_mask, q, q_c, q_mask,
  choices, choices_c, choices_mask, y, existing_emb_mat, image_emb_mat), at PINNED padded
  shapes (the reference re-derives padded shapes per batch, model_v2.py:1126-1165; results
  depend on them -- SURVEY fact 10 -- so benchmarks and data-parallel runs pin them).
* :func:`make_attention_inputs` -- encoder outputs for attention-only runs: ``hall`` ~ N(0,1)
  with masked rows exactly 0 (as ``dynamic_rnn`` leaves them), ``hq``, ``lq`` ~ N(0,1).

Length distributions follow SURVEY section 8(d): albums U[1,M]; title U[1,JXA]; description
U[0,JD]; when U[1,JT]; where U[0,JG]; photos/album U[1,JI]; photo-title U[0,JXP]; question
U[3,JQ]; choices U[1,JA]; word lengths U[1,W]; photo features unit-L2 rows; labels U{0..C-1}.
"""
import torch


def _prefix_mask(lens, J):
    return torch.arange(J).expand(*lens.shape, J) < lens.unsqueeze(-1)


def _randint(gen, lo, hi, shape):
    """Uniform integers in [lo, hi] inclusive."""
    if hi < lo:
        hi = lo
    return torch.randint(lo, hi + 1, shape, generator=gen)


def make_feed(config, seed=1234, num_examples=None, glove_vocab=300, num_photos_pool=None, dense=False):
    """Returns a dict placeholder-name -> CPU tensor.  ``num_examples`` < batch_size leaves
    all-zero padded rows at the end (model_v2.py:1103 static N; utils.py:117 short batch).
    ``dense=True`` makes every sequence full length (mask density 1)."""
    g = torch.Generator().manual_seed(seed)
    N = config.batch_size
    n_ex = N if num_examples is None else num_examples
    M, JI, JXP = config.max_num_albums, config.max_num_photos, config.max_sent_photo_title_size
    JD, JXA = config.max_sent_des_size, config.max_sent_album_title_size
    JT, JG = config.max_when_size, config.max_where_size
    JQ, JA, W = config.max_question_size, config.max_answer_size, config.max_word_size
    C = config.num_choice
    VW, VC = config.word_vocab_size, config.char_vocab_size
    V_all = VW + glove_vocab

    live = (torch.arange(N) < n_ex)
    n_alb = _randint(g, 1, M, (N,)) if not dense else torch.full((N,), M)
    n_alb = n_alb * live
    alb_live = _prefix_mask(n_alb, M)                                   # [N,M]

    def lens(lo, hi, shape, gate):
        l = _randint(g, lo, hi, shape) if not dense else torch.full(shape, hi)
        return l * gate

    at_len = lens(1, JXA, (N, M), alb_live)
    ad_len = lens(0, JD, (N, M), alb_live)
    when_len = lens(1, JT, (N, M), alb_live)
    where_len = lens(0, JG, (N, M), alb_live)
    pis_len = lens(1, JI, (N, M), alb_live)
    pis_mask = _prefix_mask(pis_len, JI)                                # [N,M,JI]
    pts_len = lens(0, JXP, (N, M, JI), pis_mask)
    q_len = lens(min(3, JQ), JQ, (N,), live)
    ch_len = lens(1, JA, (N, C), live.unsqueeze(1))

    feed = {}

    def words(name, ln, J):
        mask = _prefix_mask(ln, J)
        ids = _randint(g, 2, V_all - 1, mask.shape) * mask              # 0 = NULL, 1 = UNK (main.py:249-254)
        wl = _randint(g, 1, W, mask.shape)
        cmask = _prefix_mask(wl, W) & mask.unsqueeze(-1)
        cids = _randint(g, 2, VC - 1, cmask.shape) * cmask
        feed[name] = ids.to(torch.int32)
        feed[name + "_c"] = cids.to(torch.int32)
        feed[name + "_mask"] = mask

    words("at", at_len, JXA)
    words("ad", ad_len, JD)
    words("when", when_len, JT)
    words("where", where_len, JG)
    words("pts", pts_len, JXP)
    words("q", q_len, JQ)
    words("choices", ch_len, JA)

    P = num_photos_pool or max(8, int(pis_mask.sum().item()))
    feats = torch.randn(P, config.image_feat_dim, generator=g)
    feats = feats / feats.norm(dim=1, keepdim=True)                     # preprocess.py:220-221: pre-L2-normed
    feed["image_emb_mat"] = feats
    feed["pis"] = (_randint(g, 0, P - 1, pis_mask.shape) * pis_mask).to(torch.int32)
    feed["pis_mask"] = pis_mask
    feed["existing_emb_mat"] = torch.randn(glove_vocab, config.word_emb_size, generator=g) * 0.4
    y = torch.zeros(N, C, dtype=torch.bool)
    lab = _randint(g, 0, C - 1, (N,))
    y[torch.arange(N), lab] = True
    y &= live.unsqueeze(1)
    feed["y"] = y
    return feed


def make_attention_inputs(N, K, M, JMAX, D, JQ, seed=1234, density=1.0, dtype=torch.float32,
                          q_density=1.0):
    """``hall`` [N,K,M,JMAX,D] (masked rows exactly 0), ``hall_mask`` [N,K,M,JMAX] bool,
    ``hq`` [N,JQ,D], ``q_mask`` [N,JQ] bool, ``lq`` [N,D].  ``density`` < 1 draws a
    per-(n,k,m) prefix length so ~density of the rows are valid."""
    g = torch.Generator().manual_seed(seed)
    if density >= 1.0:
        hall_mask = torch.ones(N, K, M, JMAX, dtype=torch.bool)
    else:
        hi = max(1, int(round(2 * density * JMAX)))
        ln = torch.randint(0, min(hi, JMAX) + 1, (N, K, M), generator=g)
        hall_mask = _prefix_mask(ln, JMAX)
    if q_density >= 1.0:
        q_mask = torch.ones(N, JQ, dtype=torch.bool)
    else:
        ql = torch.randint(max(1, int(q_density * JQ)), JQ + 1, (N,), generator=g)
        q_mask = _prefix_mask(ql, JQ)
    hall = torch.randn(N, K, M, JMAX, D, generator=g, dtype=dtype) * hall_mask.unsqueeze(-1)
    hq = torch.randn(N, JQ, D, generator=g, dtype=dtype) * q_mask.unsqueeze(-1)
    lq = torch.randn(N, D, generator=g, dtype=dtype)
    return dict(hall=hall, hall_mask=hall_mask, hq=hq, q_mask=q_mask, lq=lq)
