"""CPU suite: host-side logic of the wrappers that needs no GPU (file formats, argument checks, no-fallback behaviour)."""
import numpy as np
import pytest
import torch


def test_flo_round_trip_matches_flowlib_layout(tcl, tmp_path):
    # utils/flowlib.py:33-55: 'PIEH', int32 width, int32 height, H x W x 2 float32
    rng = np.random.default_rng(0)
    flow = rng.standard_normal((7, 13, 2)).astype(np.float32)
    path = str(tmp_path / "a.flo")
    tcl.ingest.write_flo(path, flow)
    raw = open(path, "rb").read()
    assert raw[:4] == b"PIEH" and np.frombuffer(raw[4:12], np.int32).tolist() == [13, 7]
    assert len(raw) == 12 + flow.nbytes
    assert np.array_equal(tcl.ingest.read_flo(path), flow)
    open(path, "wb").write(b"XXXX" + raw[4:])
    with pytest.raises(Exception):
        tcl.ingest.read_flo(path)


def test_wrappers_refuse_cpu_tensors_instead_of_falling_back(tcl):
    x, f = torch.zeros(1, 3, 8, 8), torch.zeros(1, 2, 8, 8)
    for call in (lambda: tcl.warp(x, f), lambda: tcl.fs_warp(x, f), lambda: tcl.fbcCheckTorch(f, f),
                 lambda: tcl.gradient(f[:, 0]), lambda: tcl.temporal_error(f, f, x, x),
                 lambda: tcl.temporal_loss(torch.ones(1, 1, 8, 8), x, x, f), lambda: tcl.hwc_split(torch.zeros(1, 8, 8, 9), tcl.ingest.FC2_LAYOUT),
                 lambda: tcl.upsample_flow(f, torch.zeros(1, 576, 8, 8)), lambda: tcl.temporal_error_clip(x.repeat(2, 1, 1, 1), f, f)):
        with pytest.raises(RuntimeError):
            call()


def test_fc2_layout_is_the_reference_slicing(tcl):
    # core/data_loader.py:243-245 / learning-based/datasets.py:52-54: imgs 0:6, mask 6:7, flow 7:9
    assert tcl.ingest.FC2_LAYOUT == (("img1", 0, 3), ("img2", 3, 3), ("mask", 6, 1), ("flow", 7, 2))
    assert sum(c for _, _, c in tcl.ingest.FC2_LAYOUT) == 9
