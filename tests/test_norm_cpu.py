"""CPU: the apex-FusedRMSNorm boundary (SURVEY 8b boundary 3): constructor / parameter names / the reference's
rebinding hook.  (The kernel itself is covered by tests/test_gemm_gpu.py::test_rmsnorm* and tests/test_train_gpu.py.)"""
import types

import torch


def test_b200_rmsnorm_has_the_fused_rmsnorm_surface():
    from visionllm_b200.norm import B200RMSNorm
    m = B200RMSNorm(4096, eps=1e-6, elementwise_affine=True)       # apex.normalization.FusedRMSNorm(normalized_shape, eps, ...)
    assert list(m.state_dict()) == ["weight"] and m.weight.shape == (4096,) and m.eps == 1e-6 == m.variance_epsilon
    assert list(B200RMSNorm((8,), elementwise_affine=False).state_dict()) == []
    m2 = B200RMSNorm(normalized_shape=16)
    m2.load_state_dict({"weight": torch.full((16,), 2.0)})
    assert float(m2.weight[0]) == 2.0


def test_install_rebinds_like_the_reference_monkey_patch():
    import transformers.models.llama.modeling_llama as ml
    from visionllm_b200 import norm
    orig = ml.LlamaRMSNorm
    iv = types.ModuleType("fake_intern_vit"); iv.InternRMSNorm = object
    try:
        done = norm.install(llama=True, internvit_module=iv)
        assert "transformers.models.llama.modeling_llama.LlamaRMSNorm" in done and iv.InternRMSNorm is norm.B200RMSNorm
        layer_norm = ml.LlamaRMSNorm(64, eps=1e-5)                 # HF constructs LlamaRMSNorm(hidden_size, eps=config.rms_norm_eps)
        assert isinstance(layer_norm, norm.B200RMSNorm) and layer_norm.eps == 1e-5
        assert ml.LlamaRMSNorm(64).eps == 1e-6                     # the reference's partial(FusedRMSNorm, eps=1e-6) default
    finally:
        ml.LlamaRMSNorm = orig
