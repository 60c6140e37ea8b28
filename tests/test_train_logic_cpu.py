"""CPU: autograd structure of the training-side decoder (visionllm_b200/train.py) -- which tensor every Function saves,
in-place rotations, packed attention gradients, residuals folded into the GEMM epilogue, label shift -- against HF
`LlamaForCausalLM` autograd in fp32.  The kernels are replaced IN THIS TEST ONLY by torch fp32 stand-ins (the kernels'
own numerics are the business of tests/test_train_gpu.py); what is checked here is the host logic around them."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


@pytest.fixture()
def train_stand_ins(monkeypatch):
    import visionllm_b200.ops as ops
    import visionllm_b200.train as T
    from oracle import torch_kernels as K

    def gemm_tn(a, b, a_mn=False, b_mn=False, out_dtype=None):
        A = a.float().t() if a_mn else a.float()
        Bm = b.float().t() if b_mn else b.float()
        return A @ Bm.t()

    def rmsnorm_bwd(x2, w, dy2, eps):
        with torch.enable_grad():                          # stand-ins differentiate with autograd inside a backward
            x = x2.detach().float().requires_grad_(True)
            wf = w.detach().float().requires_grad_(True)
            (wf * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))).backward(dy2.float())
        return x.grad, wf.grad

    def attention_backward_packed(qkv5, do, scale):
        with torch.enable_grad():
            q = qkv5.detach().float().requires_grad_(True)
            B, T, _, H, D = q.shape
            s = torch.einsum("bqhd,bkhd->bhqk", q[:, :, 0], q[:, :, 1]) * scale
            s = s.masked_fill(~torch.ones(T, T, dtype=torch.bool).tril(), float("-inf"))
            o = torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1), q[:, :, 2]).reshape(B, T, H * D)
            o.backward(do.float().reshape(B, T, H * D))
        return q.grad

    class CE:
        @staticmethod
        def apply(logits, labels):
            return F.cross_entropy(logits, labels, ignore_index=-100)

    for name in ("linear", "rmsnorm", "rope_", "attention"):
        monkeypatch.setattr(ops, name, getattr(K, name))
    monkeypatch.setattr(ops, "gemm_tn", gemm_tn)
    monkeypatch.setattr(T, "rmsnorm_bwd", rmsnorm_bwd)
    monkeypatch.setattr(T, "swiglu_fwd", lambda gu: F.silu(gu[:, 0::2].float()) * gu[:, 1::2].float())

    def swiglu_bwd(gu, dh):
        with torch.enable_grad():
            g = gu.detach().float().requires_grad_(True)
            (F.silu(g[:, 0::2]) * g[:, 1::2]).backward(dh.float())
        return g.grad

    monkeypatch.setattr(T, "swiglu_bwd", swiglu_bwd)
    monkeypatch.setattr(T, "attention_backward_packed", attention_backward_packed)
    monkeypatch.setattr(T, "CrossEntropyFn", CE)


def test_train_decoder_autograd_structure_matches_hf(train_stand_ins):
    from transformers import LlamaConfig, LlamaForCausalLM
    from visionllm_b200.llama import B200LlamaForCausalLM
    from visionllm_b200.train import B200LlamaForCausalLMTrain
    cfg = LlamaConfig(hidden_size=128, intermediate_size=352, num_hidden_layers=2, num_attention_heads=2,
                      num_key_value_heads=2, vocab_size=97, rms_norm_eps=1e-5, max_position_embeddings=512,
                      attn_implementation="eager")
    torch.manual_seed(0)
    hf = LlamaForCausalLM(cfg).float().eval()
    B, T = 2, 256
    gen = torch.Generator().manual_seed(3)
    emb = torch.randn(B, T, 128, generator=gen) * 0.5
    labels = torch.randint(0, 97, (B, T), generator=gen)
    labels[:, :100] = -100

    e = emb.clone().requires_grad_(True)
    out = hf(inputs_embeds=e, attention_mask=torch.ones(B, T, dtype=torch.long))
    ref_loss = F.cross_entropy(out.logits[:, :-1].reshape(-1, 97), labels[:, 1:].reshape(-1), ignore_index=-100)
    ref_loss.backward()
    ref_grads = {n: p.grad.clone() for n, p in hf.named_parameters() if p.grad is not None}   # embed_tokens is bypassed
    ref_demb = e.grad.clone()

    lm = B200LlamaForCausalLM(cfg)
    lm.load_state_dict(hf.state_dict(), strict=True)
    lm = lm.float()
    tr = B200LlamaForCausalLMTrain(lm)
    for _ in range(2):                                     # twice: nothing stale is carried between steps
        for p in lm.parameters():
            p.grad = None
        e2 = emb.clone().requires_grad_(True)
        loss, logits, hidden = tr(e2, labels)
        loss.backward()
        assert abs(float(loss) - float(ref_loss)) < 1e-4 * abs(float(ref_loss))
        assert (logits - out.logits).abs().max() < 1e-3
        assert (e2.grad - ref_demb).abs().max() <= 1e-3 * ref_demb.abs().max() + 1e-6
        got = dict(lm.named_parameters())
        assert set(ref_grads) <= set(got)
        for n, g in ref_grads.items():
            assert got[n].grad is not None, n
            assert (got[n].grad.float() - g).abs().max() <= 2e-3 * g.abs().max() + 1e-6, n
