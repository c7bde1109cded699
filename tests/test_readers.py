"""Format readers (VDB-224 + Blosc/LZ4, Radiance HDR, BMP, EXR) against the committed fixtures and the
files' own metadata -- the only pins these formats offer."""
import os

import numpy as np
import pytest

import vpt_b200 as V
from vpt_b200.scene import load_vdb_grid, load_bmp_rbg, load_exr_rgb, load_hdr, find_asset


def test_dragon_vdb_matches_file_metadata_and_survey_numbers():
    dens, meta = load_vdb_grid(find_asset("dragon.vdb"), "density")
    assert dens.shape == (31, 49, 70) and meta["dim"] == (70, 49, 31)
    assert meta["bbox_min"] == (16, 1, 35) and meta["bbox_max"] == (85, 49, 65)
    assert meta["active_voxels"] == 19660 and meta["leaf_count"] == 131           # file_voxel_count metadata, checked inside the reader too
    assert meta["max_value"] == 1.0 and meta["min_density"] == np.float32(np.finfo(np.float32).eps)   # quirk Q10
    assert abs(float(dens.mean()) - 0.0958042) < 1e-6 and abs(float((dens != 0).mean()) - 0.185) < 1e-3
    assert abs(meta["voxel_size"] - 0.1) < 1e-7
    X = meta["xform"]
    assert np.allclose(np.diag(X), [0.1, 0.1, 0.1, 1.0]) and np.count_nonzero(X) == 4


def test_vdb_missing_grid_and_bad_file(tmp_path):
    assert load_vdb_grid(find_asset("dragon.vdb"), "no_such_grid") is None
    bad = tmp_path / "bad.vdb"; bad.write_bytes(b"not a vdb file at all, definitely")
    with pytest.raises(V.VptError):
        load_vdb_grid(str(bad), "density")
    trunc = tmp_path / "trunc.vdb"; trunc.write_bytes(open(find_asset("dragon.vdb"), "rb").read()[:5000])
    with pytest.raises(V.VptError):
        load_vdb_grid(str(trunc), "density")


@pytest.mark.skipif(not os.path.exists("/root/reference/assets/fireball.vdb"), reason="large reference assets not on this box")
def test_large_reference_vdbs_decode_to_their_metadata():
    dens, m = load_vdb_grid("/root/reference/assets/fireball.vdb", "density")
    assert m["dim"] == (239, 257, 234) and m["active_voxels"] == 4901516 and abs(m["max_value"] - 12.4696) < 1e-3
    heat, mh = load_vdb_grid("/root/reference/assets/fireball.vdb", "heat")
    assert mh["dim"] == (270, 288, 268) and mh["active_voxels"] == 20839680 and abs(mh["background"] + 0.01) < 1e-6
    cd, mc = load_vdb_grid("/root/reference/assets/colored_smoke.vdb", "Cd")
    assert mc["channels"] == 3 and cd.shape[-1] == 3 and mc["active_voxels"] == 1676306
    d2, m2 = load_vdb_grid("/root/reference/assets/dragon_with_xform.vdb", "density")
    assert m2["active_voxels"] == 156161 and m2["dim"] == (141, 99, 63)


def test_blue_noise_bmp_channel_swap():
    bn = load_bmp_rbg(find_asset("BN0.bmp"))
    assert bn.shape == (256, 256, 3)
    # first pixel of BN0.bmp is R=26 G=70 B=108: x = R/255, y = B/255, z = G/255 (quirk Q16)
    assert np.allclose(bn[0, 0], [26 / 255.0, 108 / 255.0, 70 / 255.0])
    assert 0.45 < bn.mean() < 0.55


def test_exr_luts():
    bb = load_exr_rgb(find_asset("blackbody_texture.exr"))
    assert bb.shape == (1, 256, 3)
    assert np.allclose(bb[0, 0], [5.9604645e-06, 1.0430813e-05, 2.6345253e-05], rtol=1e-6)
    assert np.allclose(bb[0, -1], [0.9711914, 0.8720703, 0.72265625])
    dc = load_exr_rgb(find_asset("density_color_texture2.exr"))
    assert dc.shape == (1, 256, 3) and np.all(dc == 1.0)


def test_hdr_rgbe_rule(tmp_path):
    # two scanlines: one flat (non-RLE because width < 8), values follow (m + 0.5) * 2^(e - 136)
    p = tmp_path / "t.hdr"
    px = bytes([128, 64, 32, 129,  0, 0, 0, 0,  255, 255, 255, 128,  1, 2, 3, 120])
    p.write_bytes(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 2 +X 2\n" + px)
    img = load_hdr(str(p))
    assert img.shape == (2, 2, 4)
    assert np.allclose(img[0, 0, :3], [(128 + 0.5) * 2.0 ** -7, (64 + 0.5) * 2.0 ** -7, (32 + 0.5) * 2.0 ** -7])
    assert np.all(img[0, 1, :3] == 0) and np.all(img[..., 3] == 0)
    assert np.allclose(img[1, 0, :3], [(255.5) * 2.0 ** -8] * 3)
    hdr = find_asset("Barce_Rooftop_C_3k.hdr")
    if hdr:
        big = load_hdr(hdr)
        assert big.shape == (1500, 3000, 4) and abs(float(big[..., :3].mean()) - 0.245057) < 1e-4


def test_ins_scene_file_reader(tmp_path):
    """`.ins` grammar of the reference's read_instance_file (main.cpp:980-1040): volumes form."""
    from vpt_b200.scene import load_ins
    p = tmp_path / "two.ins"
    p.write_text("2\n./assets/dragon.vdb\n3\n0 0 0 0 0 0 1 1\n1.5 -2 3e0 0.0 0.7071068 0 0.7071068 0.5\n-4 5 6 1 0 0 0 2\r\n"
                 "clouds/cumulus 01.vdb\n1\n10 20 30 0 0 0 1 0.25\n")
    d = load_ins(str(p))
    assert d["kind"] == "volumes" and [f["path"] for f in d["files"]] == ["./assets/dragon.vdb", "clouds/cumulus 01.vdb"]
    assert [len(f["instances"]) for f in d["files"]] == [3, 1]
    pos, quat, scale = d["files"][0]["instances"][1]
    assert pos == (1.5, -2.0, 3.0) and quat == (0.0, 0.7071068, 0.0, 0.7071068) and scale == 0.5
    assert d["files"][1]["instances"][0] == ((10.0, 20.0, 30.0), (0.0, 0.0, 0.0, 1.0), 0.25)


def test_ins_light_file_reader_and_errors(tmp_path):
    from vpt_b200.scene import load_ins
    p = tmp_path / "lights.ins"
    p.write_text("light\n2\n9 6 2 1 0.8 0.6 40\n-2 3 8 0.5 0.7 1 25\n")
    d = load_ins(str(p))
    assert d == {"kind": "lights", "lights": [((9.0, 6.0, 2.0), (1.0, 0.8, 0.6), 40.0), ((-2.0, 3.0, 8.0), (0.5, 0.7, 1.0), 25.0)]}
    bad = tmp_path / "short.ins"
    bad.write_text("1\nfoo.vdb\n2\n0 0 0 0 0 0 1 1\n1 2 3\n")                 # second record has too few numbers
    with pytest.raises(V.VptError, match="8 numbers"):
        load_ins(str(bad))
    with pytest.raises(V.VptError, match="cannot open"):
        load_ins(str(tmp_path / "missing.ins"))
    empty = tmp_path / "empty.ins"; empty.write_text("")
    with pytest.raises(V.VptError, match="empty file"):
        load_ins(str(empty))
