"""ORACLE (test infrastructure): fp32 CPU restatement of the reference's ViT / LLM / projector layers in
plain torch ops.  Floating-point path, so the oracle is torch fp32 (not C).  Each function cites what it follows.

Pinned by tests/test_oracle_modules_cpu.py against
  * tests/golden/mod_internvit_small.npz (outputs of the reference's own InternVisionModel), and
  * the installed HF LlamaDecoderLayer (the reference's LLM is third-party transformers==4.34.0,
    requirements.txt:23 -- not vendored; same arithmetic in the installed 5.5).
Used by bench.py's cpu_baseline / --impl reference legs as the reference's CPU path of the forward.
"""
import math

import torch
import torch.nn.functional as F


def rmsnorm(x, w, eps):
    # internvit/modeling_intern_vit.py:38-44 == apex manual_rms_norm (fused_layer_norm.py:16-29)
    v = x.float().pow(2).mean(-1, keepdim=True)
    return w * (x.float() * torch.rsqrt(v + eps)).to(x.dtype)


def internvit_embeddings(px, sd, patch):
    # modeling_intern_vit.py:82-90
    pe = F.conv2d(px, sd["embeddings.patch_embedding.weight"], sd["embeddings.patch_embedding.bias"], stride=patch)
    pe = pe.flatten(2).transpose(1, 2)
    cls = sd["embeddings.class_embedding"].expand(px.shape[0], 1, -1)
    return torch.cat([cls, pe], 1) + sd["embeddings.position_embedding"]


def internvit_layer(x, sd, pre, heads, eps, qk_norm=True):
    # modeling_intern_vit.py:126-143 (_naive_attn), :175-179, :198-210
    B, N, C = x.shape
    h = rmsnorm(x, sd[pre + "norm1.weight"], eps)
    qkv = F.linear(h, sd[pre + "attn.qkv.weight"], sd.get(pre + "attn.qkv.bias"))
    qkv = qkv.reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    if qk_norm:
        q = rmsnorm(q.transpose(1, 2).flatten(-2, -1), sd[pre + "attn.q_norm.weight"], eps).view(B, N, heads, -1).transpose(1, 2)
        k = rmsnorm(k.transpose(1, 2).flatten(-2, -1), sd[pre + "attn.k_norm.weight"], eps).view(B, N, heads, -1).transpose(1, 2)
    att = ((q * (C // heads) ** -0.5) @ k.transpose(-2, -1)).softmax(-1)
    a = (att @ v).transpose(1, 2).reshape(B, N, C)
    a = F.linear(a, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])
    x = x + a * sd[pre + "ls1"]
    h = rmsnorm(x, sd[pre + "norm2.weight"], eps)
    h = F.linear(F.gelu(F.linear(h, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])),
                 sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
    return x + h * sd[pre + "ls2"]


def clip_layer(x, sd, pre, heads, eps=1e-5):
    # transformers/models/clip/modeling_clip.py CLIPEncoderLayer (the released-7B preset's tower, constant.py / train.py:350-352):
    # pre-LN, q/k/v/out projections with bias, softmax(q k^T / sqrt(d)), quick-GELU MLP
    B, N, C = x.shape
    h = F.layer_norm(x, (C,), sd[pre + "layer_norm1.weight"], sd[pre + "layer_norm1.bias"], eps)
    q, k, v = (F.linear(h, sd[pre + f"self_attn.{n}_proj.weight"], sd[pre + f"self_attn.{n}_proj.bias"])
               .view(B, N, heads, C // heads).transpose(1, 2) for n in "qkv")
    att = ((q * (C // heads) ** -0.5) @ k.transpose(-2, -1)).softmax(-1)
    a = (att @ v).transpose(1, 2).reshape(B, N, C)
    x = x + F.linear(a, sd[pre + "self_attn.out_proj.weight"], sd[pre + "self_attn.out_proj.bias"])
    h = F.layer_norm(x, (C,), sd[pre + "layer_norm2.weight"], sd[pre + "layer_norm2.bias"], eps)
    h = F.linear(h, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])
    return x + F.linear(h * torch.sigmoid(1.702 * h), sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])


def internvit_forward(px, sd, layers, heads, patch, eps=1e-6):
    x = internvit_embeddings(px, sd, patch)
    states = [x]
    for i in range(layers):
        x = internvit_layer(x, sd, f"encoder.layers.{i}.", heads, eps)
        states.append(x)
    return states


def rotate_half(x):
    return torch.cat((-x[..., x.shape[-1] // 2:], x[..., : x.shape[-1] // 2]), -1)


def llama_layer(x, sd, pre, heads, eps, theta=10000.0):
    # HF LlamaDecoderLayer: RMSNorm -> q,k,v -> rotate-half RoPE -> causal softmax(fp32) -> o_proj -> +res
    #                       -> RMSNorm -> down(silu(gate) * up) -> +res
    B, T, H = x.shape
    D = H // heads
    h = rmsnorm(x, sd[pre + "input_layernorm.weight"], eps)
    q = F.linear(h, sd[pre + "self_attn.q_proj.weight"]).view(B, T, heads, D).transpose(1, 2)
    k = F.linear(h, sd[pre + "self_attn.k_proj.weight"]).view(B, T, heads, D).transpose(1, 2)
    v = F.linear(h, sd[pre + "self_attn.v_proj.weight"]).view(B, T, heads, D).transpose(1, 2)
    inv = 1.0 / (theta ** (torch.arange(0, D, 2).float() / D))
    fr = torch.outer(torch.arange(T).float(), inv)
    emb = torch.cat((fr, fr), -1)
    cos, sin = emb.cos()[None, None].to(x.dtype), emb.sin()[None, None].to(x.dtype)
    q, k = q * cos + rotate_half(q) * sin, k * cos + rotate_half(k) * sin
    s = (q @ k.transpose(-1, -2)) / math.sqrt(D)
    s = s + torch.full((T, T), float("-inf")).triu(1)
    a = (s.float().softmax(-1).to(x.dtype) @ v).transpose(1, 2).reshape(B, T, H)
    x = x + F.linear(a, sd[pre + "self_attn.o_proj.weight"])
    h = rmsnorm(x, sd[pre + "post_attention_layernorm.weight"], eps)
    m = F.linear(F.silu(F.linear(h, sd[pre + "mlp.gate_proj.weight"])) * F.linear(h, sd[pre + "mlp.up_proj.weight"]),
                 sd[pre + "mlp.down_proj.weight"])
    return x + m
