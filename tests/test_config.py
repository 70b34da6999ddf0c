"""CPU: the CLI boundary -- every flag of the reference's main.py:28-147 (frozen in
tests/golden/main_flags.json by tools/make_golden.py) is accepted with the same type and
default; the post-parse rules of main.py:151-208 hold."""
import json
import os

import pytest

from fvta_memexqa_b200.config import build_parser, get_args, fvta_config

GOLD = os.path.join(os.path.dirname(__file__), "golden", "main_flags.json")


def test_every_reference_flag_same_type_and_default():
    ref = json.load(open(GOLD))
    assert len(ref) == 64                      # 2 positional + 62 flags
    actions = {a.option_strings[0] if a.option_strings else a.dest: a for a in build_parser()._actions}
    for name, kw in ref.items():
        assert name in actions, name
        a = actions[name]
        if kw.get("action") == "store_true":
            assert a.nargs == 0 and a.const is True and a.default is False, name
        else:
            if "type" in kw:
                assert a.type.__name__ == kw["type"], name
            if "default" in kw:
                assert a.default == kw["default"], name
    assert "--num_choice" in actions and actions["--num_choice"].default == 4   # model_v2.py:381


def test_post_parse_rules(tmp_path):
    base = [str(tmp_path / "prepro"), str(tmp_path / "out")]
    a = get_args(base + ["--is_test", "--load_best", "--modelname", "fvta", "--runId", "3", "--keep_prob", "0.5"])
    assert a.load and a.num_epochs == 1 and a.keep_prob == 1.0           # main.py:182-191
    assert a.outpath.endswith(os.path.join("fvta", "03"))
    assert a.val_path.endswith("test_best") and os.path.isdir(a.val_path)
    assert os.path.isdir(a.save_dir) and os.path.isdir(a.save_dir_best)
    with pytest.raises(AssertionError):
        get_args(base)                                                   # neither --is_train nor --is_test
    with pytest.raises(AssertionError):
        get_args(base + ["--is_test", "--is_pack_model"])                # main.py:151-154
    # argparse prefix matching is relied on by the README (--use_bidirect, README.MD:125)
    assert get_args(base + ["--is_train", "--use_bidirect"]).use_bidirection


def test_fvta_config_is_the_readme_command():
    c = fvta_config()
    assert (c.hidden_size, c.simiMatrix, c.warp_type, c.image_feat_dim, c.image_trans_dim) == (50, 2, 5, 2537, 100)
    assert c.use_3d and c.use_time_warp and c.add_tanh and c.use_question_att and c.use_char
    assert (c.max_num_albums, c.max_num_photos, c.max_question_size) == (4, 8, 25)
