"""Headless driver (rtgs/__main__.py): the reference's flags and defaults (__main__.py:41-85), the display
conversion, and (GPU) an end-to-end run on tests/data/test.ply checked against the float64 oracle."""
import numpy as np
import pytest

from rtgs.__main__ import build_parser, main, to_display


def test_flags_and_defaults_match_the_reference():
    a = build_parser().parse_args([])
    assert a.res == (960, 540) and a.fov == 90 and a.sample == 1 and a.depth == 16 and a.bvh == 1024 and a.scale == 1
    a = build_parser().parse_args(["-o", "x.ply", "-r", "256,128", "-f", "60", "-s", "2", "-d", "8", "-v", "64",
                                   "--scale", "30"])
    assert str(a.open) == "x.ply" and a.res == (256, 128) and a.fov == 60 and a.sample == 2 and a.depth == 8
    assert a.bvh == 64 and a.scale == 30


def test_display_conversion_is_bottom_left_origin():
    disp = np.zeros((4, 3, 3), np.float32)      # (W,H,3), [i,j] = column, row from the bottom
    disp[1, 0] = (1.0, 0.5, 2.0)                # column 1, bottom row
    img = to_display(disp)
    assert img.shape == (3, 4, 3)
    assert tuple(img[2, 1]) == (255, 128, 255)  # bottom row of a top-left-origin image, clipped
    assert img.sum() == 255 + 128 + 255


def test_missing_scene_is_an_error(capsys):
    assert main([]) == 2


@pytest.mark.gpu
def test_cli_renders_the_reference_scene(tmp_path, test_ply):
    from oracle import ref_numpy as O
    from rtgs.orbit import focal_from_fov, orbit_pose
    from rtgs.ply import read_ply
    rc = main(["-o", str(test_ply), "-r", "128,96", "-f", "90", "-s", "2", "--scale", "30", "--t-cut", "0",
               "--out", str(tmp_path), "--format", "both"])
    assert rc == 0
    disp = np.load(tmp_path / "frame.npy")
    pos, rot = orbit_pose(0.0, np.pi / 2, 1.0)
    f = focal_from_fov(96, 90.0)
    gs = O.GaussianSet(**O.activate(read_ply(test_ply), 30.0))
    ref = O.render(gs, O.CameraParams(np.asarray(pos), np.asarray(rot), 128, 96, (f, f)), depth=16)["rgb"]
    assert np.abs(disp - ref).max() <= 1e-3      # two identical samples averaged
    from PIL import Image
    png = np.asarray(Image.open(tmp_path / "frame.png"))
    assert png.shape == (96, 128, 3) and png.max() > 0
    # 4-view orbit sweep
    assert main(["-o", str(test_ply), "-r", "64,48", "--scale", "30", "--views", "4", "--out", str(tmp_path / "sweep"),
                 "--format", "npy"]) == 0
    assert sorted(p.name for p in (tmp_path / "sweep").iterdir()) == [f"view_{k:03d}.npy" for k in range(4)]
    # (one sample per view: the sweep went through the pipelined RayTracer.sweep) every view is its own pose
    f = focal_from_fov(48, 90.0)
    for k in (1, 3):
        pos, rot = orbit_pose(2 * np.pi * k / 4, np.pi / 2, 1.0)
        ref = O.render(gs, O.CameraParams(np.asarray(pos), np.asarray(rot), 64, 48, (f, f)), depth=16)["rgb"]
        assert np.abs(np.load(tmp_path / "sweep" / f"view_{k:03d}.npy") - ref).max() <= 1e-3
