"""CPU: the numpy restatements (oracle/) against the golden vectors produced by
the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import codec_numpy as OC
from oracle import mse_numpy as OM

import _golden_checks as GC

IMAGES = ["pe", "torax", "synth16_257x301", "synth12_300x200", "synth8_129x70", "sat12_96x160"]


@pytest.mark.parametrize("name", IMAGES)
def test_entropy_mi_split(golden, golden_images, name):
    GC.check_entropy_and_split(OC, golden_images[name], golden["images"][name])


@pytest.mark.parametrize("name", IMAGES)
@pytest.mark.parametrize("case_idx", range(6))
def test_lsb_cases(golden, golden_images, name, case_idx):
    GC.check_lsb_case(OC, OM, golden_images[name], golden["images"][name]["lsb_cases"][case_idx])


def test_appendix_b_headline(golden):
    """SURVEY.md Appendix B rows, as a guard on the golden file itself."""
    pe = golden["images"]["pe"]["lsb_cases"][0]
    assert pe["segments_lengths"] == [163, 91, 40, 10] and pe["segment_indices"] == [2, 1, 3, 0]
    assert pe["stego_sha"].startswith("cc8b6f0a85c4c757") and pe["bitmaps_sha"].startswith("89d89bd07d3d7bac")
    assert pe["px_changed"] == 141 and pe["mse"] == 0.002895355224609375 and pe["max_range"] == 836.0
    tx = golden["images"]["torax"]["lsb_cases"][0]
    assert tx["segments_lengths"] == [197, 86, 21] and tx["segment_indices"] == [1, 0, 2]
    assert tx["stego_sha"].startswith("108fec83c15c1cc4") and tx["px_changed"] == 174


def test_segment_plan(golden):
    for key, rec in golden["segments"].items():
        s, total = (int(v) for v in key.split(":"))
        sizes, order = OC.segment_plan(s, total)
        assert sizes == rec["sizes"] and order == rec["order"]
        segs = OC._segments(np.zeros(total, np.uint8), sizes, order)
        assert [len(x) for x in segs] == rec["seg_lens"]


def test_scalars(golden):
    sc = golden["scalars"]
    m, r = OM.calcular_mse([[10, 20], [30, 40]], [[10, 20], [30, 41]])
    assert [float(m), float(r)] == sc["mse_norm_small"] == [0.21875, 41.0]
    from codec_tcc_b200.synth import synth_image
    a = synth_image(120, 90, 4095, 21)
    b = a.copy(); b[5, 7] += 900; b[60:70, 10:50] ^= 3
    assert [float(v) for v in OM.calcular_mse(a, b)] == sc["mse_norm_synth"]
    assert float(OM.calcular_ssim_simples(a, b)) == sc["ssim_norm_synth"]
    assert [float(v) for v in OM.calcular_mse(a, a)] == sc["mse_same"]
    assert OM.calcular_psnr(0.0, 4095) == float("inf") == sc["psnr_zero"]
    assert float(OM.calcular_psnr(1.0)) == sc["psnr_default"]
    assert float(OM.calcular_psnr(0.37, 4095)) == sc["psnr_4095"]
    assert OC.message_to_bits("Olá, DICOM ✓") == sc["message_bits_hex"]
    with pytest.raises(ValueError):
        OM.calcular_mse(np.zeros((2, 3)), np.zeros((3, 2)))


def test_general_mutual_information_and_float_metrics(golden):
    """Round-2 golden entries (produced by the unmodified reference): planes that are not bit planes of the
    image, and metric inputs that are not integer-valued 8/16-bit data."""
    from oracle import mse_numpy as OM
    for name, plane, img in GC.general_mi_cases():
        assert float(OC.calculate_mutual_information(plane, img)) == golden["scalars"]["mi_general"][name], name
    for name, x, y in GC.float_metric_cases():
        want = golden["scalars"]["float_metrics"][name]
        m, r = OM.calcular_mse(x, y)
        assert float(m) == want["mse"] and float(r) == want["max_range"], name
        assert float(OM.calcular_ssim_simples(x, y)) == want["ssim"], name
