"""CPU: pin the oracle against the reference's own outputs (tests/golden/*.npz,
produced by tests/golden/gen_golden.py running the reference's
multi_scale_deformable_attn_pytorch) and against the mmcv unit-test tolerances
(mmcv/tests/test_ops/test_ms_deformable_attn.py:72-134)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import msda_oracle as O

KNOWN_SEED3 = [0.0018993784157779181, 0.004602827532968805, 0.004671175247309776, 0.004384399819001662,
               0.0037950971737622935, 0.002512764199421532, 0.0018444261512603004, 0.003634679248037905]


def test_golden_file_is_the_survey_known_answer(golden_dir):
    g = np.load(os.path.join(golden_dir, "msda_mmcv_seed3.npz"))
    assert np.allclose(g["out_f64"].flatten(), KNOWN_SEED3, rtol=0, atol=1e-18)


def test_kernel_oracle_mmcv_seed3_fp64(golden_dir):
    g = np.load(os.path.join(golden_dir, "msda_mmcv_seed3.npz"))
    out = O.forward_kernel_semantics(g["value"].astype(np.float64), g["shapes"], g["lsi"],
                                     g["loc"].astype(np.float64), g["attw"].astype(np.float64))
    ref = g["out_f64"]
    # the reference's own CUDA-vs-pytorch fp64 bounds (test_ms_deformable_attn.py:99-102)
    assert np.abs(out - ref).max() < 1e-18
    assert (np.abs(out - ref) / np.abs(ref)).max() < 1e-15


def test_kernel_oracle_mmcv_seed3_fp32(golden_dir):
    g = np.load(os.path.join(golden_dir, "msda_mmcv_seed3.npz"))
    out = O.forward_kernel_semantics(g["value"], g["shapes"], g["lsi"], g["loc"], g["attw"])
    ref = g["out_f32"]
    assert np.allclose(out, ref, rtol=1e-2, atol=1e-3)          # :129
    assert np.abs(out - ref).max() < 1e-9                        # :133
    assert (np.abs(out - ref) / np.abs(ref)).max() < 1e-6        # :134


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(
    os.path.join(os.path.dirname(__file__), "golden", "msda_ref_*.npz"))))
def test_kernel_oracle_vs_reference_fp64(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    out = O.forward_kernel_semantics(g["value"].astype(np.float64), g["shapes"], g["lsi"],
                                     g["loc"].astype(np.float64), g["attw"].astype(np.float64))
    ref = g["out_f64"]
    # kernel formula (loc*W-0.5) vs grid_sample formula (((2loc-1)+1)*W-1)/2 differ by rounding only.
    # Exception, faithfully kept: the reference kernel calls floorf() even for double (kernel.cuh:22-23),
    # so a double h_im within 1e-7 below an integer is floored one cell too high and the bilinear weights
    # extrapolate by ~1e-7 -- only the pixel-centre case (fp32 refs widened to fp64) triggers it.
    tol = 5e-6 if "pixel" in name else 1e-12
    assert np.abs(out - ref).max() < tol
    out32 = O.forward_kernel_semantics(g["value"], g["shapes"], g["lsi"], g["loc"], g["attw"])
    # fp32: a sample within ~1e-6 of a cell edge may pick the neighbouring cell; bilinear is continuous,
    # so the value moves by O(1e-6 * |v|)
    assert np.abs(out32 - ref).max() < 2e-5


@pytest.mark.parametrize("name", ["msda_ref_d32_npot.npz", "msda_ref_d32_pixel.npz"])
def test_grid_sample_restatement_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    out = O.forward_grid_sample(torch.from_numpy(g["value"]).double(), torch.from_numpy(g["shapes"]),
                                torch.from_numpy(g["loc"]).double(), torch.from_numpy(g["attw"]).double())
    assert np.abs(out.numpy() - g["out_f64"]).max() < 1e-13


def test_sample_indices_properties():
    rng = np.random.default_rng(0)
    shapes = np.array([[13, 17], [100, 37]], dtype=np.int64)
    loc = (rng.random((2, 50, 3, 2, 4, 2), dtype=np.float32) * 1.3 - 0.15).astype(np.float32)
    idx = O.sample_indices(shapes, loc)
    H = shapes[:, 0][None, None, None, :, None]
    W = shapes[:, 1][None, None, None, :, None]
    valid = (idx[..., 2] & 1) == 1
    assert valid.any() and (~valid).any()
    # in-range samples have h_low in [-1, H-1], w_low in [-1, W-1]
    assert ((idx[..., 0] >= -1) & (idx[..., 0] <= H - 1))[valid].all()
    assert ((idx[..., 1] >= -1) & (idx[..., 1] <= W - 1))[valid].all()
    # corner bits agree with the bounds
    ll = (idx[..., 0] >= 0) & (idx[..., 1] >= 0)
    assert (((idx[..., 2] >> 1) & 1).astype(bool) == ll)[valid].all()
    assert (idx[~valid] == 0).all()


def test_pixel_centre_non_power_of_two():
    """SURVEY Appendix A adversarial case: refs at (i+0.5)/W with W=100 -- the
    un-contracted formula must be used (h_low follows fl(fl(loc*W) - 0.5))."""
    W = 100
    xs = ((np.arange(W, dtype=np.float32) + np.float32(0.5)) / np.float32(W)).astype(np.float32)
    loc = np.stack([xs, xs], -1).reshape(1, W, 1, 1, 1, 2)
    idx = O.sample_indices(np.array([[W, W]], dtype=np.int64), loc)
    prod = (xs * np.float32(W)).astype(np.float32)
    expect = np.floor((prod - np.float32(0.5)).astype(np.float32)).astype(np.int32)
    assert (idx[0, :, 0, 0, 0, 1] == expect).all()
    assert (idx[0, :, 0, 0, 0, 0] == expect).all()


@pytest.mark.parametrize("name", ["dcnv3_ref_testpy.npz", "dcnv3_ref_c32_s2.npz", "dcnv3_ref_c32_k5.npz"])
def test_dcnv3_oracle_vs_reference(golden_dir, name):
    """oracle/dcnv3_oracle.c vs outputs of the reference's own dcnv3_core_pytorch (fp64); the reference's own
    fp32 check is allclose(rtol=1e-2, atol=1e-3) (ops_dcnv3/test.py:79)."""
    from oracle import dcnv3_oracle as DO
    g = np.load(os.path.join(golden_dir, name))
    p = [int(x) for x in g["params"]]
    out = DO.forward(g["input"], g["offset"], g["mask"], *p[:8], p[8], p[9], float(g["offset_scale"]))
    ref = g["out_f64"]
    assert np.allclose(out, ref, rtol=1e-2, atol=1e-3)
    assert np.abs(out - ref).max() < 1e-7


@pytest.mark.parametrize("name", ["msda_bwd_d32.npz", "msda_bwd_d4.npz", "msda_bwd_d30.npz", "msda_bwd_d71.npz"])
def test_backward_oracle_vs_reference_autograd(golden_dir, name):
    """C backward restatement vs autograd through the reference's own pytorch function (fp64)."""
    g = np.load(os.path.join(golden_dir, name))
    gv, gl, gw = O.backward_kernel_semantics(g["value"], g["shapes"], g["lsi"], g["loc"], g["attw"], g["grad_out"])
    assert np.abs(gv - g["grad_value"]).max() < 1e-10
    assert np.abs(gl - g["grad_loc"]).max() < 1e-9 * max(1.0, np.abs(g["grad_loc"]).max())
    assert np.abs(gw - g["grad_attw"]).max() < 1e-10


@pytest.mark.parametrize("name", ["dcnv3_bwd_testpy.npz", "dcnv3_bwd_c32_s2.npz"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_dcnv3_backward_oracle_vs_reference_autograd(golden_dir, name, dtype):
    """oracle_dcnv3_backward_* vs fp64 autograd through the reference's own dcnv3_core_pytorch.  The reference builds
    its sampling grid with float32 linspace (dcnv3_func.py:85-101), so even its fp64 run carries ~1e-7 noise."""
    from oracle import dcnv3_oracle as DO
    g = np.load(os.path.join(golden_dir, name))
    p = [int(x) for x in g["params"]]
    got = DO.backward(g["input"], g["offset"], g["mask"], g["grad_out"], *p, float(g["offset_scale"]), dtype=dtype)
    for a, key in zip(got, ("grad_input", "grad_offset", "grad_mask")):
        assert np.abs(a - g[key]).max() <= 5e-6 * np.abs(g[key]).max(), key
