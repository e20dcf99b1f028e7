"""get_target_keys semantics (what /root/reference/tests/test_utils.py:12-127 pins): substring include,
anchored-regex include, exclude wins."""
from src.utils.dtype import str_to_dtype
from src.utils.state_dict import RegexMatch, get_target_keys
from src.utils.tensor import remove_orig_mod_prefix

import pytest
import torch

KEYS = [
    "model.layers.0.attn.q_proj",
    "model.layers.0.attn.k_proj",
    "model.layers.0.mlp.fc1",
    "model.layers.1.attn.q_proj",
    "model.layers.1.mlp.fc1",
    "head",
]


def test_substring_include():
    assert sorted(get_target_keys(["attn"], [], KEYS)) == [k for k in sorted(KEYS) if "attn" in k]
    assert get_target_keys([], [], KEYS) == []
    assert sorted(get_target_keys(["q_proj", "head"], [], KEYS)) == ["head", "model.layers.0.attn.q_proj", "model.layers.1.attn.q_proj"]


def test_regex_is_anchored_at_start():
    assert sorted(get_target_keys([RegexMatch(regex=r".*layers\.\d+\.mlp")], [], KEYS)) == [
        "model.layers.0.mlp.fc1",
        "model.layers.1.mlp.fc1",
    ]
    # re.match: no leading ".*" -> must match from the first character
    assert get_target_keys([RegexMatch(regex=r"layers\.\d+")], [], KEYS) == []
    assert get_target_keys([RegexMatch(regex=r"head$")], [], KEYS) == ["head"]


def test_exclude_wins():
    got = get_target_keys(["model."], ["layers.1", RegexMatch(regex=r".*k_proj")], KEYS)
    assert sorted(got) == ["model.layers.0.attn.q_proj", "model.layers.0.mlp.fc1"]
    assert get_target_keys(["attn"], ["attn"], KEYS) == []


def test_shipped_auraflow_regex_matches_nothing_without_attn():
    """the shipped configs/auraflow/lora.yml regex omits '.attn' and matches no module."""
    keys = [f"denoiser.single_layers.{i}.attn.w1q" for i in range(3)]
    assert get_target_keys([RegexMatch(regex=r".*single_layers\.\d+\.w1[qkvo]")], [], keys) == []
    assert len(get_target_keys([RegexMatch(regex=r".*single_layers\.\d+\.attn\.w1[qkvo]")], [], keys)) == 3


def test_regexmatch_callable():
    assert RegexMatch(regex=r"a.c")("abc") and not RegexMatch(regex=r"a.c")("xabc")


def test_dtype_names_and_prefix():
    assert str_to_dtype("BF16") is torch.bfloat16 and str_to_dtype("float") is torch.float32
    assert str_to_dtype("fp16") is torch.float16
    with pytest.raises(ValueError):
        str_to_dtype("int8")
    assert remove_orig_mod_prefix("_orig_mod.a._orig_mod.b") == "a._orig_mod.b"
