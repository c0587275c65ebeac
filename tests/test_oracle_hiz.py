"""CPU: the oracle's far-HiZ builder (lane-by-lane walk of nvhiz-update.comp.glsl) against an independent texel-level
restatement of the same dispatch schedule, plus known answers."""
import numpy as np
import pytest

from oracle.oracle_binding import Oracle
from vk_tessellated_clusters_b200 import scenes as S


def hiz_reference_numpy(depth: np.ndarray):
    """Texel-level restatement of NVHizVK::cmdUpdateHiz (nvhiz_vk.cpp:484-594): per dispatch, level i over the whole
    (8-aligned) dispatch extent from clamped 2x2 source footprints, levels i+1 / i+2 as 2x2 maxima of the level below
    (what the shuffles compute), stores outside a level dropped."""
    h, w = depth.shape
    size, mips, _, _, _, _ = S.hiz_info(w, h)
    levels = [np.zeros((max(1, size >> l), max(1, size >> l)), np.float32) for l in range(mips)]
    in_w, in_h = w, h
    sub_w, sub_h = (w + 1) // 2, (h + 1) // 2
    for i in range(0, mips, 3):
        src = depth if i == 0 else levels[i - 1]
        sub_w, sub_h = (sub_w + 7) // 8 * 8, (sub_h + 7) // 8 * 8
        ow, oh = (sub_w + 7) // 8 * 8, (sub_h + 7) // 8 * 8
        cx = np.minimum(np.arange(ow) * 2, in_w - 2)
        cy = np.minimum(np.arange(oh) * 2, in_h - 2)
        pad = np.zeros((max(src.shape[0], cy.max() + 2), max(src.shape[1], cx.max() + 2)), np.float32)  # fetches outside the level read 0
        pad[:src.shape[0], :src.shape[1]] = src
        a = pad[np.ix_(cy, cx)]
        b = pad[np.ix_(cy, cx + 1)]
        c = pad[np.ix_(cy + 1, cx)]
        d = pad[np.ix_(cy + 1, cx + 1)]
        cur = np.maximum(np.maximum(np.maximum(a, b), c), d)
        for l in range(3):
            if i + l >= mips:
                break
            if l > 0:
                cur = np.maximum(np.maximum(cur[0::2, 0::2], cur[0::2, 1::2]), np.maximum(cur[1::2, 0::2], cur[1::2, 1::2]))
            n = levels[i + l].shape[0]
            hh, ww = min(n, cur.shape[0]), min(n, cur.shape[1])
            levels[i + l][:hh, :ww] = cur[:hh, :ww]
        for _ in range(3):
            sub_w, sub_h = (sub_w + 1) // 2, (sub_h + 1) // 2
        sub_w, sub_h = max(sub_w, 1), max(sub_h, 1)
        in_w, in_h = sub_w * 2, sub_h * 2
    return np.concatenate([l.reshape(-1) for l in levels]), size, mips


SIZES = [(2, 2), (16, 16), (37, 23), (23, 37), (64, 48), (200, 120), (130, 258), (255, 257), (640, 360)]


@pytest.mark.parametrize("w,h", SIZES)
def test_oracle_matches_texel_level_restatement(w, h):
    rng = np.random.default_rng(2342 + w * 1000 + h)
    depth = rng.random((h, w), dtype=np.float32)
    o = Oracle()
    o.update_hiz(depth)
    pyr, size, mips = o.get_hiz()
    ref, rsize, rmips = hiz_reference_numpy(depth)
    assert (size, mips) == (rsize, rmips)
    assert pyr.tobytes() == ref.tobytes()
    o.close()


def test_shape_and_factors_match_reference_formulas():
    o = Oracle()
    for w, h in [(3840, 2160), (1920, 1080), (37, 23), (1024, 1024)]:  # (usedW - 2 wraps for images under 4 texels, as in the reference)
        size, mips, factors, smax = o.hiz_info(w, h)
        esize, emips, uw, uh, efac, esmax = S.hiz_info(w, h)
        assert (size, mips, smax) == (esize, emips, esmax)
        assert factors.tobytes() == efac.tobytes()
    # README / nvhiz_vk.cpp:290-308 known answer: 3840x2160 -> 2048^2, 12 levels, used 1920x1080
    assert o.hiz_info(3840, 2160)[:2] == (2048, 12)
    o.close()


def test_pow2_image_equals_plain_max_mip_chain():
    """For a square pow2 depth image nothing is clamped: level l is the plain 2x2-max chain of the image."""
    rng = np.random.default_rng(7)
    depth = rng.random((128, 128), dtype=np.float32)
    o = Oracle()
    o.update_hiz(depth)
    pyr, size, mips = o.get_hiz()
    assert (size, mips) == (64, 7)
    half = np.maximum(np.maximum(depth[0::2, 0::2], depth[0::2, 1::2]), np.maximum(depth[1::2, 0::2], depth[1::2, 1::2]))
    chain, n = S.make_hiz_pyramid(half)
    assert n == mips and pyr.tobytes() == chain.tobytes()
    assert pyr[-1] == depth.max()
    o.close()


def test_update_replaces_set_hiz_and_rejects_tiny_images():
    o = Oracle()
    o.set_hiz(np.ones(5, np.float32), 2, 2)
    o.update_hiz(np.full((16, 16), 0.25, np.float32))
    pyr, size, mips = o.get_hiz()
    assert (size, mips) == (8, 4) and np.all(pyr == 0.25)
    with pytest.raises(Exception):
        o.update_hiz(np.zeros((1, 8), np.float32))
    o.close()


@pytest.mark.parametrize("w,h", SIZES + [(1023, 511), (1920, 1080)])
def test_oracle_matches_reference_hiz_shader(w, h):
    """The oracle's pyramid against the REFERENCE'S OWN nvhiz-update.comp.glsl, compiled for the host and run by the SIMT
    emulator (oracle/ref/, two pipelines NV_HIZ_IS_FIRST 1/0, dispatch schedule of NVHizVK::cmdUpdateHiz): bit-identical."""
    from oracle import ref_binding
    from vk_tessellated_clusters_b200 import api

    try:
        ref = ref_binding.ReferenceShaders(api.Config(), True)
    except SystemExit as e:
        pytest.skip(str(e))
    depth = np.random.default_rng(w * 7919 + h).random((h, w), dtype=np.float32)
    o = Oracle()
    for _ in range(2):  # second update reuses the pyramid storage
        o.update_hiz(depth)
        ref.update_hiz(depth)
    a, sa, ma = ref.get_hiz()
    b, sb, mb = o.get_hiz()
    assert (sa, ma) == (sb, mb)
    assert a.tobytes() == b.tobytes()
