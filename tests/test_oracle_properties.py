"""Domain properties of the oracle's definitions (tie rule, shard additivity, cost, edge cases)."""
import numpy as np

from hybridquantization_b200 import synth


def test_tie_rule_lowest_index_wins(oracle):
    # OptimizedConvolution.cl:186 replaces only on strict '<': duplicated colours -> first index
    img = synth.synth_image(32, 8, 7)
    pal = synth.synth_palettes(1, 6)
    pal[0, 4] = pal[0, 1]
    pal[0, 5] = pal[0, 0]
    for space in (oracle.SPACE_LAB, oracle.SPACE_SRGB):
        r = oracle.assign_reduce(img, pal, space, want_idx=True, threads=1)
        assert r["counts"][0, 4] == 0 and r["counts"][0, 5] == 0
        assert set(np.unique(r["idx"][0])) <= {0, 1, 2, 3}


def test_exact_palette_colour_has_zero_error(oracle):
    img = np.array([[[10, 200, 30], [10, 200, 30], [250, 0, 9]]], np.uint8)
    unit, _ = oracle.image_planes(img)
    pal = np.zeros((1, 2, 4), np.float32)
    pal[0, 0, :3] = unit[:, 0]
    pal[0, 1, :3] = unit[:, 2]
    for space in (oracle.SPACE_LAB, oracle.SPACE_SRGB):
        r = oracle.assign_reduce(img, pal, space, want_idx=True)
        assert r["err_fx"][0] == 0
        assert r["idx"][0].tolist() == [0, 0, 1]
        assert r["counts"][0].tolist() == [2, 1]


def test_row_shards_add_up_exactly(oracle):
    img = synth.synth_image(40, 30, 11, smooth=True)
    pal = synth.synth_palettes(2, 9)
    whole = oracle.assign_reduce(img, pal, threads=1)
    for cuts in ([0, 30], [0, 7, 30], [0, 1, 2, 29, 30], [0, 0, 30]):
        err = np.zeros(2, np.int64); cnt = np.zeros((2, 9), np.uint64); sums = np.zeros((2, 9, 3), np.int64)
        for a, b in zip(cuts[:-1], cuts[1:]):
            if a == b:
                continue
            r = oracle.assign_reduce(img[a:b], pal, threads=2)
            err += r["err_fx"]; cnt += r["counts"]; sums += r["sums_fx"]
        assert np.array_equal(err, whole["err_fx"]) and np.array_equal(cnt, whole["counts"]) and np.array_equal(sums, whole["sums_fx"])


def test_cost_is_mean_plus_penalty(oracle):
    counts = np.array([5, 0, 3, 0], np.uint64)
    c = oracle.cost(3 << 24, counts, 8, 2.0)  # sum dE = 3.0 over 8 pixels, two unused colours
    assert c == 3.0 / 8 + 4.0
    assert oracle.cost(0, np.array([1], np.uint64), 1, 2.0) == 0.0


def test_counts_sum_to_pixels_and_centroids_inside_gamut(oracle):
    img = synth.synth_image(50, 20, 3)
    pal = synth.synth_palettes(1, 16)
    r = oracle.assign_reduce(img, pal)
    assert int(r["counts"].sum()) == 1000
    used = r["counts"][0] > 0
    cent = r["sums_fx"][0][used] / 16777216.0 / r["counts"][0][used, None]
    assert (cent[:, 0] >= 0).all() and (cent[:, 0] <= 100.001).all()


def test_single_colour_palette_and_single_pixel(oracle):
    img = np.array([[[1, 2, 3]]], np.uint8)
    pal = synth.synth_palettes(1, 1)
    r = oracle.assign_reduce(img, pal, want_idx=True)
    assert r["counts"][0, 0] == 1 and r["idx"][0, 0] == 0 and r["err_fx"][0] > 0


def test_quantize_outputs_palette_colours(oracle):
    img = synth.synth_image(16, 16, 5)
    pal = synth.synth_palettes(1, 8)[0]
    q = oracle.quantize(img, pal)
    assert np.array_equal(q["f32"], pal[q["idx"]])
    want = (pal[q["idx"], :3] * np.float32(255.0) + np.float32(0.5)).astype(np.int32).astype(np.uint8)
    assert np.array_equal(q["rgb"], want)
    # idempotence: every output colour is a palette colour, re-quantising keeps the indices' colours
    q2 = oracle.quantize(q["rgb"].reshape(16, 16, 3), pal)
    assert np.array_equal(q2["rgb"], oracle.quantize(q2["rgb"].reshape(16, 16, 3), pal)["rgb"])
