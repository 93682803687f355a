"""Host side of the drop-in boundary (no GPU): settings.config parser, the
environment generator (glibc rand stream) and the image loader."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import load_golden

SHIPPED = """# Mode (1 = random environment mode (default), 2 = loaded image mode)
mode=1

ncols=1000
nrows=1000
nb_of_obstacles=15
minWidth=100
maxWidth=200
minHeight=100
maxHeight=200
randomSeed=0
seedValue=25
imagePath=path\\images\\maze_6.png
start={50,50}
end={990,990}
max_iter=250
visibilityThreshold=0.25
lightStrength=1
timer=1
saveResults=1
saveCameFrom=1
saveLightSources=1
saveGlobalVisibility=1
saveLocalVisibility=1
saveVisibilityField=1
silent=1
ballRadius=15
"""


@pytest.fixture(scope="module")
def vhp():
    import visibility_heuristic_path_planner_b200 as m
    m.load_library()
    return m


def parse(vhp, tmp_path, text):
    f = tmp_path / "settings.config"
    f.write_text(text)
    cfg = vhp.Config()
    lib = vhp.load_library()
    lib.vhp_config_parse.argtypes = [C.c_char_p, C.POINTER(vhp.Config)]
    st = lib.vhp_config_parse(str(f).encode(), C.byref(cfg))
    return st, cfg


def test_parse_shipped_config(vhp, tmp_path):
    st, c = parse(vhp, tmp_path, SHIPPED)
    assert st == 0
    assert (c.mode, c.ncols, c.nrows, c.nb_of_obstacles) == (1, 1000, 1000, 15)
    assert (c.min_width, c.max_width, c.min_height, c.max_height) == (100, 200, 100, 200)
    assert (c.random_seed, c.seed_value) == (0, 25)
    assert (c.start_x, c.start_y, c.end_x, c.end_y) == (50, 50, 990, 990)
    assert c.max_iter == 250 and c.visibility_threshold == 0.25 and c.ball_radius == 15
    assert c.image_path == b"path\\images\\maze_6.png"


def test_defaults_and_errors(vhp, tmp_path):
    st, c = parse(vhp, tmp_path, "# only comments\n\nsilent = true\nbogus=3\n")
    assert st == 0                       # unknown keys only warn
    assert (c.ncols, c.nrows, c.nb_of_obstacles, c.max_iter) == (100, 100, 10, 100)
    assert c.visibility_threshold == 0.5 and c.random_seed == 1 and c.ball_radius == 5
    assert parse(vhp, tmp_path, "silent=1\ntimer=maybe\n")[0] == -1
    assert parse(vhp, tmp_path, "silent=1\nvisibilityThreshold=1.5\n")[0] == -1
    assert parse(vhp, tmp_path, "silent=1\nncols=-4\n")[0] == -1
    st, c = parse(vhp, tmp_path, "silent=1\nmode=7\nstart={3;4}\n")
    assert st == 0 and c.mode == 1 and (c.start_x, c.start_y) == (0, 0)
    lib = vhp.load_library()
    assert lib.vhp_config_parse(b"/nonexistent/settings.config", C.byref(vhp.Config())) == -4


def test_environment_generator_matches_reference_stream(vhp, oracle):
    lib = vhp.load_library()
    lib.vhp_environment_generate.argtypes = [C.POINTER(vhp.Config), C.c_void_p, C.POINTER(C.c_int64)]
    for (nx, ny, nobs, mnw, mxw, mnh, mxh, seed) in [(101, 101, 10, 10, 20, 10, 20, 2),
                                                      (1000, 1000, 15, 100, 200, 100, 200, 1),
                                                      (64, 200, 30, 1, 5, 2, 40, 9)]:
        cfg = vhp.Config()
        lib.vhp_config_default(C.byref(cfg))
        cfg.ncols, cfg.nrows, cfg.nb_of_obstacles = nx, ny, nobs
        cfg.min_width, cfg.max_width, cfg.min_height, cfg.max_height = mnw, mxw, mnh, mxh
        cfg.random_seed, cfg.seed_value, cfg.silent = 0, seed, 1
        occ = np.zeros((ny, nx), dtype=np.uint8)
        used = C.c_int64(-1)
        assert lib.vhp_environment_generate(C.byref(cfg), occ.ctypes.data, C.byref(used)) == 0
        assert used.value == seed
        assert np.array_equal(occ, oracle.generate_environment(nx, ny, nobs, mnw, mxw, mnh, mxh, seed))


def test_image_loader(vhp, tmp_path):
    from PIL import Image
    lib = vhp.load_library()
    lib.vhp_environment_load_image.argtypes = [C.c_char_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    g = load_golden("maze5.npz")
    ny, nx = map(int, g["shape"])
    occ = np.unpackbits(g["occ_bits"])[: ny * nx].reshape(ny, nx)
    rgba = np.zeros((ny, nx, 4), dtype=np.uint8)
    rgba[..., 0] = np.where(occ, 255, 37)       # red channel == 255 -> free
    rgba[..., 1] = 90; rgba[..., 3] = 255
    variants = {"rgba.png": Image.fromarray(rgba, "RGBA"), "rgb.png": Image.fromarray(rgba[..., :3], "RGB"),
                "gray.png": Image.fromarray(np.where(occ, 255, 0).astype(np.uint8), "L"),
                "pal.png": Image.fromarray(rgba[..., :3], "RGB").convert("P", palette=Image.ADAPTIVE, colors=4),
                "gray.pgm": Image.fromarray(np.where(occ, 255, 10).astype(np.uint8), "L"),
                "rgb.ppm": Image.fromarray(rgba[..., :3], "RGB")}
    for name, im in variants.items():
        p = str(tmp_path / name)
        im.save(p)
        w, h = C.c_int(0), C.c_int(0)
        assert lib.vhp_environment_load_image(p.encode(), None, C.byref(w), C.byref(h)) == 0, name
        assert (w.value, h.value) == (nx, ny)
        out = np.zeros((ny, nx), dtype=np.uint8)
        assert lib.vhp_environment_load_image(p.encode(), out.ctypes.data, C.byref(w), C.byref(h)) == 0
        expect = (np.array(Image.open(p).convert("RGBA"))[..., 0] == 255).astype(np.uint8)
        assert np.array_equal(out, expect), name
        if name != "pal.png":
            assert np.array_equal(out, occ), name
    assert lib.vhp_environment_load_image(b"/nonexistent.png", None, C.byref(w), C.byref(h)) == -4
