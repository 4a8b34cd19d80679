"""tuatara_b200.convert: the reference's weight files (TorchScript archives, tuatara.cpp:333,:423) and wrapped /
pickled checkpoints export to the same .ttw bytes as a direct state_dict export."""
import torch

from tuatara_b200 import convert, weights


def _same_file(a, b):
    return open(a, "rb").read() == open(b, "rb").read()


def test_torchscript_archives_convert(oracle_models, tmp_path):
    craft, parseq = oracle_models
    # CRAFT as the reference ships it: a traced TorchScript module (file name of tuatara.cpp:333)
    traced = torch.jit.trace(craft, torch.zeros(1, 3, 64, 64), check_trace=False)
    traced.save(str(tmp_path / convert.CRAFT_FILE))
    # PARSeq wrapped the way upstream's Lightning system does (keys "model.*"), pickled under the reference's file name
    torch.save({"state_dict": {"model." + k: v for k, v in parseq.state_dict().items()}}, tmp_path / convert.PARSEQ_FILE)
    assert convert.main(["--weights-dir", str(tmp_path)]) == 0
    weights.export_craft(craft.state_dict(), tmp_path / "craft_direct.ttw")
    weights.export_parseq(parseq.state_dict(), tmp_path / "parseq_direct.ttw")
    assert _same_file(tmp_path / "craft.ttw", tmp_path / "craft_direct.ttw")
    assert _same_file(tmp_path / "parseq.ttw", tmp_path / "parseq_direct.ttw")


def test_dataparallel_prefix_and_bad_file(oracle_models, tmp_path):
    craft, _ = oracle_models
    sd = convert.normalise({"module." + k: v for k, v in craft.state_dict().items()}, "basenet.slice1.0.weight")
    assert set(sd) == set(craft.state_dict())
    try:
        convert.normalise({"foo.weight": torch.zeros(1)}, "basenet.slice1.0.weight")
    except KeyError as e:
        assert "not the expected architecture" in str(e)
    else:
        raise AssertionError("a foreign state_dict must be rejected")


def test_parseq_torchscript_archive_and_pickle_policy(oracle_models, tmp_path):
    """PARSeq as the reference ships it (a TorchScript archive under the file name of tuatara.cpp:423) converts to the
    same bytes as the direct export; a pickled nn.Module (arbitrary code on load) is refused unless explicitly trusted."""
    import pytest

    craft, parseq = oracle_models
    torch.jit.trace(craft, torch.zeros(1, 3, 64, 64), check_trace=False).save(str(tmp_path / convert.CRAFT_FILE))
    torch.jit.trace(parseq, torch.zeros(1, 3, 32, 128), check_trace=False).save(str(tmp_path / convert.PARSEQ_FILE))
    assert convert.main(["--weights-dir", str(tmp_path)]) == 0
    weights.export_parseq(parseq.state_dict(), tmp_path / "parseq_direct.ttw")
    assert _same_file(tmp_path / "parseq.ttw", tmp_path / "parseq_direct.ttw")
    # a whole pickled module: weights_only refuses it
    mod = tmp_path / "module.pkl"
    torch.save(craft, mod)
    with pytest.raises(ValueError, match="unsafe"):
        convert.load_state_dict(mod)
    assert set(convert.load_state_dict(mod, unsafe_pickle=True)) == set(craft.state_dict())
