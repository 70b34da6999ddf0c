"""Command-line flags and the ``config`` namespace of the reference, kept flag-for-flag.

Mirrors the interface of the reference's ``main.py:25-210`` (``get_args``): same flag names,
types and defaults, the same positional arguments, the same derived output paths and the same
post-parse mode rules.  One flag is added: ``--num_choice`` (default 4) -- the reference's
``model_v2.Model`` reads ``config.num_choice`` (model_v2.py:381) but its ``main.py`` never
defines it.

``update_config``-style fields (``max_*``, vocab sizes; reference ``utils.py:294-396``) are
written onto the same namespace by :func:`set_shapes` for synthetic data, or by
``utils.update_config`` for real data.
"""
import argparse
import os

# (flag, kwargs) in the order of main.py:28-147
_FLAGS = [
    ("--modelname", dict(type=str, default="memoryqa")),
    ("--runId", dict(type=int, default=0, help="used for run the same model multiple times")),
    ("--use_3d", dict(action="store_true", help="use 3D tensor model")),
    ("--dmnplus", dict(action="store_true", help="DMN+ model exp")),
    ("--dmnplus_num_hops", dict(type=int, default=3, help="dmn hop times")),
    ("--mcb", dict(action="store_true", help="MCB attention model exp")),
    ("--mcb_outdim", dict(type=int, default=16000, help="MCB attention model exp")),
    ("--no_photo", dict(action="store_true", default=False, help="text-only experiment")),
    ("--load", dict(action="store_true", default=False, help="whether to load existing model")),
    ("--load_best", dict(action="store_true", default=False, help="whether to load the best model")),
    ("--is_train", dict(action="store_true", default=False, help="training mode, ")),
    ("--is_test", dict(action="store_true", default=False, help="testing mode, otherwise test mode")),
    ("--get_yp", dict(action="store_true", default=False, help="testing mode, whether to save all yp")),
    ("--is_test_on_val", dict(action="store_true", default=False, help="test on validation set")),
    ("--is_save_weights", dict(action="store_true", default=False, help="whether to save model weights to val_path")),
    ("--is_save_vis", dict(action="store_true", default=False,
                           help="whether to save each layer output for visualization during testing, will save into val_path")),
    ("--showspecs", dict(action="store_true", default=False, help="show the meta of the data and then exit")),
    ("--save_period", dict(type=int, default=200, help="num steps to save model and eval")),
    ("--val_path", dict(type=str, default="", help="path to store the eval file[for testing]")),
    ("--batch_size", dict(type=int, default=20)),
    ("--val_num_batches", dict(type=int, default=100, help="eval during training, get how many batch in train/val to eval")),
    ("--num_epochs", dict(type=int, default=20)),
    ("--word_count_thres", dict(default=2, type=int, help="word count threshold")),
    ("--char_count_thres", dict(default=10, type=int, help="char count threshold")),
    ("--sent_album_title_size_thres", dict(default=10, type=int, help="max sentence word count for album_title")),
    ("--sent_photo_title_size_thres", dict(default=8, type=int, help="max sentence word count for photo_title")),
    ("--sent_des_size_thres", dict(default=10, type=int, help="max sentence word count for album_description")),
    ("--sent_when_size_thres", dict(default=4, type=int, help="max sentence word count for album_when")),
    ("--sent_where_size_thres", dict(default=4, type=int, help="max sentence word count for album_where")),
    ("--answer_size_thres", dict(default=5, type=int, help="answer word count")),
    ("--question_size_thres", dict(default=25, type=int, help="max question word count")),
    ("--word_size_thres", dict(default=16, type=int, help="max word character count")),
    ("--num_photos_thres", dict(default=10, type=int, help="maximum photo number per album")),
    ("--num_albums_thres", dict(default=8, type=int, help="maximum album number")),
    ("--hidden_size", dict(type=int, default=100)),
    ("--use_char", dict(default=False, action="store_true", help="use character CNN embeding")),
    ("--char_emb_size", dict(default=8, type=int, help="char-CNN channel size")),
    ("--char_out_size", dict(default=100, type=int, help="char-CNN output size for each word")),
    ("--concat", dict(default=False, action="store_true", help="For simple lstm exp, whether to concat all modality")),
    ("--keep_prob", dict(default=1.0, type=float, help="1.0 - drop out rate;remember to set it to 1.0 in eval")),
    ("--wd", dict(default=None, type=float, help="l2 weight decay loss, 0.002 is a good number, default not applied")),
    ("--image_feat_dim", dict(default=3048, type=int, help="image feature length")),
    ("--use_image_trans", dict(default=False, action="store_true", help="use image transform from image_feat_dim to image_tran_dim")),
    ("--image_trans_dim", dict(default=300, type=int, help="image transformed feature length")),
    ("--init_lr", dict(default=0.5, type=float, help="Start learning rate")),
    ("--use_ml_att", dict(default=False, action="store_true", help="use multi-level attention")),
    ("--use_tgif_ml_att", dict(default=False, action="store_true", help="use TGIF temporal attention")),
    ("--use_mm_att", dict(default=False, action="store_true", help="use multi-modal attention")),
    ("--use_bidirection", dict(default=False, action="store_true", help="bidirectional attention in all attention layer")),
    ("--use_time_warp", dict(default=False, action="store_true", help="whether to use time warping to get new context")),
    ("--warp_type", dict(default=1, type=int, help="time warping type,1:all,2:current, 3:past,4:future,5:past-future")),
    ("--use_time_warp_att", dict(default=False, action="store_true", help="whether to use time warping to get attention logit")),
    ("--add_tanh", dict(default=False, action="store_true", help="whether to add tanh activation on attention layer")),
    ("--simiMatrix", dict(default=1, type=int,
                          help="similarity matrixed used in attention layer, 1:(h,q,h*q), 2:((h-q)^2,h*q), 3:1 & 2, 4: cosine")),
    ("--use_direct_links", dict(action="store_true", help="whether to use direct links to attend all info and add to g1")),
    ("--direct_links_only", dict(action="store_true", help="")),
    ("--use_question_att", dict(default=False, action="store_true", help="use info reqresentation to attend question")),
    ("--use_choices_att", dict(default=False, action="store_true", help="use question reqresentation to attend choices")),
    ("--use_eu_output", dict(action="store_true", help="whether to add eu similarity in the output layer")),
    ("--is_pack_model", dict(action="store_true", default=False,
                             help="with is_test, this will pack the model to a path instead of testing")),
    ("--pack_model_path", dict(type=str, default=None, help="path to save model")),
    ("--pack_model_note", dict(type=str, default=None, help="leave a note for this packed model for future reference")),
]

# Not in the reference's main.py; see module docstring.
_ADDED_FLAGS = [
    ("--num_choice", dict(type=int, default=4, help="number of answer choices (model_v2.py:381 reads it)")),
    ("--seed", dict(type=int, default=1234, help="[added] global seed for init / shuffling / answer placement")),
    ("--synthetic", dict(type=int, default=0, help="[added] generate this many synthetic MemexQA-shaped QAs "
                                                   "instead of reading prepropath (no dataset offline)")),
]


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("prepropath", type=str)
    parser.add_argument("outbasepath", type=str, help="full path will be outbasepath/modelname/runId")
    for flag, kw in _FLAGS + _ADDED_FLAGS:
        parser.add_argument(flag, **kw)
    return parser


def _mkdir(path):
    if not os.path.exists(path):
        os.makedirs(path)


def get_args(argv=None, make_dirs=True):
    """Parse ``argv`` like the reference's ``get_args`` (main.py:25-210) and apply its
    post-parse rules.  Model-module selection (main.py:156-165) is done by
    ``fvta_memexqa_b200.model.select_get_model``; only ``--use_3d`` (model_v2, the FVTA
    path) has a B200 implementation."""
    args = build_parser().parse_args(argv)

    if args.is_pack_model:  # main.py:151-154
        assert args.is_test, "use pack model with is_test"
        assert args.pack_model_path is not None, "please provide where pack model to"
        assert args.pack_model_note is not None, "please provide some note for the packed model"

    # main.py:167-179
    args.outpath = os.path.join(args.outbasepath, args.modelname, str(args.runId).zfill(2))
    args.save_dir = os.path.join(args.outpath, "save")
    args.save_dir_model = os.path.join(args.save_dir, "save")
    args.save_dir_best = os.path.join(args.outpath, "best")
    args.save_dir_best_model = os.path.join(args.save_dir_best, "save-best")
    args.write_self_sum = True
    args.self_summary_path = os.path.join(args.outpath, "train_sum.txt")
    if make_dirs:
        for p in (args.outpath, args.save_dir, args.save_dir_best):
            _mkdir(p)

    if args.load_best:  # main.py:182-183
        args.load = True

    if not args.is_train:  # main.py:186-208
        assert args.is_test, "if not train, please use is_test flag"
        args.load = True
        args.num_epochs = 1
        args.keep_prob = 1.0
        if args.val_path == "":
            args.val_path = os.path.join(args.outpath, "test_best" if args.load_best else "test")
        print("test result will be in %s" % args.val_path)
        args.vis_path = os.path.join(args.val_path, "vis")
        args.weights_path = os.path.join(args.val_path, "weights")
        if make_dirs:
            _mkdir(args.val_path)
            if args.is_save_vis:
                _mkdir(args.vis_path)
                print("visualization output will be in  %s" % args.vis_path)
            if args.is_save_weights:
                _mkdir(args.weights_path)
                print("model weights will be in %s" % args.weights_path)
    return args


def default_config(**overrides):
    """A config namespace with every flag at the reference default, then ``overrides``."""
    args = build_parser().parse_args(["", ""])
    for k, v in overrides.items():
        setattr(args, k, v)
    return args


# README.MD:222-226 -- the FVTA training command (minus run-control flags).
FVTA_FLAGS = dict(
    use_3d=True, hidden_size=50, image_feat_dim=2537, use_image_trans=True, image_trans_dim=100,
    use_char=True, char_emb_size=100, char_out_size=100, add_tanh=True, simiMatrix=2,
    use_question_att=True, use_time_warp=True, warp_type=5,
    num_albums_thres=4, num_photos_thres=8, sent_des_size_thres=10,
    sent_album_title_size_thres=8, sent_photo_title_size_thres=8, init_lr=0.5)


def set_shapes(config, M, JI, JXP, JD, JXA, JT, JG, JQ, JA, W, word_vocab_size, char_vocab_size,
               word_emb_size=100):
    """Write the data-derived fields ``update_config`` (utils.py:294-396) would write."""
    config.max_num_albums = M
    config.max_num_photos = JI
    config.max_sent_photo_title_size = JXP
    config.max_sent_des_size = JD
    config.max_sent_album_title_size = JXA
    config.max_when_size = JT
    config.max_where_size = JG
    config.max_question_size = JQ
    config.max_answer_size = JA
    config.max_word_size = W
    config.word_vocab_size = word_vocab_size
    config.char_vocab_size = char_vocab_size
    config.word_emb_size = word_emb_size
    return config


def fvta_config(batch_size=32, **overrides):
    """The FVTA configuration of README.MD:222-226 at the C2 shape."""
    kw = dict(FVTA_FLAGS)
    kw["batch_size"] = batch_size
    kw["is_train"] = True
    kw.update(overrides)
    cfg = default_config(**kw)
    set_shapes(cfg, M=cfg.num_albums_thres, JI=cfg.num_photos_thres, JXP=cfg.sent_photo_title_size_thres,
               JD=cfg.sent_des_size_thres, JXA=cfg.sent_album_title_size_thres, JT=cfg.sent_when_size_thres,
               JG=cfg.sent_where_size_thres, JQ=cfg.question_size_thres, JA=cfg.answer_size_thres,
               W=cfg.word_size_thres, word_vocab_size=400, char_vocab_size=80)
    return cfg
