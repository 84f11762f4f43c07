"""oracle/depth_oracle.py against the golden vectors the unmodified reference wrote (tests/golden/make_golden_depth.py). CPU only."""
import numpy as np

from conftest import golden
from oracle import depth_oracle as D


def test_byte_functions_match_reference_golden():
    g = golden("depth_mip")
    for k in range(int(g["b_count"])):
        img, depth, n = g["b_img_%d" % k], g["b_depth_%d" % k], int(g["b_n_%d" % k])
        planes = np.stack(D.generate_mip_layers(img, depth, n), 0)
        assert np.array_equal(planes, g["b_planes_%d" % k]), k
        assert np.array_equal(D.reconstruct_mip_image(list(g["b_styl_%d" % k]), depth, n), g["b_merged_%d" % k]), k


def test_bin_edges_are_shared_and_sums_wrap():
    """a pixel exactly on an inner edge belongs to two planes (util.py:30 is inclusive on both sides); the uint8 sum wraps"""
    depth = np.array([[0, 2, 4]], dtype=np.uint8)
    img = np.full((1, 3, 3), 200, dtype=np.uint8)
    planes = D.generate_mip_layers(img, depth, 2)
    assert planes[0][0, 1, 0] == 200 and planes[1][0, 1, 0] == 200
    merged = D.reconstruct_mip_image([img, img], depth, 2)
    assert merged[0, 1, 0] == (400 % 256) and merged[0, 0, 0] == 200 and merged[0, 2, 0] == 200


def test_style_mip_matches_reference_golden(vgg_weights):
    ws, bs = vgg_weights
    g = golden("depth_mip")
    n, steps = int(g["mip_n"]), int(g["mip_num_steps"])
    planes = D.generate_mip_layers(g["mip_content"], g["mip_depth"], n)
    assert np.array_equal(np.stack(planes, 0), g["mip_planes"])
    final, stylized, results = D.style_mip(ws, bs, g["mip_content"], g["mip_style"], g["mip_depth"], n, steps)
    for i in range(n):
        ours = np.array([l[0] for l in results[i].losses])
        ref = g["mip_losses"][i]
        assert len(ours) == len(ref) == 20 * (steps // 20 + 1)
        assert abs(ours[0] - ref[0]) <= 1e-5 * abs(ref[0]), i
        assert np.abs(ours - ref).max() <= 1e-2 * np.abs(ref).max(), i
        diff = np.abs(stylized[i].astype(int) - g["mip_stylized"][i].astype(int))
        assert diff.max() <= 8 and diff.mean() < 0.5, (i, diff.max(), diff.mean())
    diff = np.abs(final.astype(int) - g["mip_final"].astype(int))
    assert diff.mean() < 0.5
