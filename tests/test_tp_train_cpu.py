"""CPU, world_size 2 over gloo: host logic of the tensor-parallel fwd+bwd (visionllm_b200/tp_train.py; BASELINE cfg 5) --
which rows / columns of every weight a rank owns, where the two all-reduces of a layer sit, which gradients end up
sharded vs replicated.  The op set is injected: here torch-native differentiable fp32 ops; on the GPU train.py's kernel
Functions (tests/test_train_gpu.py checks those against autograd).  Expected values: the installed HF
`LlamaForCausalLM` (the third-party module the reference trains, modeling_visionllmv2.py:143, 741-757) unsharded, fp32."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def torch_fns():
    def rms(x, w, eps):
        return w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))

    def rope(qkv2, cos, sin, heads, D):
        t = qkv2.shape[0]
        v = qkv2[:, :heads * D].reshape(t, heads, D)
        rot = torch.cat((-v[..., D // 2:], v[..., :D // 2]), -1)
        return torch.cat(((v * cos[:, None, :] + rot * sin[:, None, :]).reshape(t, heads * D), qkv2[:, heads * D:]), 1)

    def attention(q, k, v, scale):
        B, T, Hh, D = q.shape
        s = torch.einsum("bqhd,bkhd->bhqk", q, k) * scale
        s = s.masked_fill(~torch.ones(T, T, dtype=torch.bool).tril(), float("-inf"))
        return torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1), v).reshape(B, T, Hh * D)

    return {"linear": F.linear, "linear_f32": F.linear, "rmsnorm": rms, "rope": rope, "attention": attention,
            "swiglu": lambda gu: F.silu(gu[:, 0::2]) * gu[:, 1::2],
            "ce": lambda logits, labels: F.cross_entropy(logits, labels, ignore_index=-100)}


def _config():
    from transformers import LlamaConfig
    return LlamaConfig(hidden_size=64, intermediate_size=96, num_hidden_layers=2, num_attention_heads=4,
                       num_key_value_heads=4, vocab_size=50, rms_norm_eps=1e-5, attn_implementation="eager")


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from transformers import LlamaForCausalLM
    from visionllm_b200 import tp_train
    cfg = _config()
    torch.manual_seed(0)
    ref = LlamaForCausalLM(cfg).float()
    B, T, H = 2, 8, cfg.hidden_size
    gen = torch.Generator().manual_seed(1)
    emb = torch.randn(B, T, H, generator=gen)
    labels = torch.randint(0, 50, (B, T), generator=gen)
    labels[:, :3] = -100
    e_ref = emb.clone().requires_grad_(True)
    out = ref(inputs_embeds=e_ref, attention_mask=torch.ones(B, T, dtype=torch.int64), labels=labels)
    out.loss.backward()
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    shards = tp_train.shard_for_training(sd, cfg, rank, world)
    model = tp_train.TPLlamaTrain(cfg, shards, fns=torch_fns())
    e = emb.clone().requires_grad_(True)
    loss, logits, _ = model(e, labels)
    loss.backward()
    g = {n: p.grad for n, p in ref.named_parameters()}
    D, il = H // cfg.num_attention_heads, cfg.intermediate_size // world
    ql = cfg.num_attention_heads // world * D
    errs = {"loss": abs(float(loss) - float(out.loss)), "logits": float((logits - out.logits).abs().max()),
            "dembeds": float((e.grad - e_ref.grad).abs().max())}
    for i, ly in enumerate(shards["layers"]):
        p = f"model.layers.{i}."
        want_qkv = torch.cat([g[p + f"self_attn.{n}_proj.weight"][rank * ql:(rank + 1) * ql] for n in ("q", "k", "v")], 0)
        gg, uu = g[p + "mlp.gate_proj.weight"][rank * il:(rank + 1) * il], g[p + "mlp.up_proj.weight"][rank * il:(rank + 1) * il]
        errs[f"l{i}"] = max(float((ly["wqkv"].grad - want_qkv).abs().max()),
                            float((ly["wo"].grad - g[p + "self_attn.o_proj.weight"][:, rank * ql:(rank + 1) * ql]).abs().max()),
                            float((ly["wgu"].grad - torch.stack([gg, uu], 1).reshape(2 * il, H)).abs().max()),
                            float((ly["wdown"].grad - g[p + "mlp.down_proj.weight"][:, rank * il:(rank + 1) * il]).abs().max()),
                            float((ly["ln1"].grad - g[p + "input_layernorm.weight"]).abs().max()),      # replicated: FULL gradient
                            float((ly["ln2"].grad - g[p + "post_attention_layernorm.weight"]).abs().max()))
    errs["top"] = max(float((shards["top"]["final_norm"].grad - g["model.norm.weight"]).abs().max()),
                      float((shards["top"]["lm_head"].grad - g["lm_head.weight"]).abs().max()))
    q.put((rank, errs, tuple(shards["layers"][0]["wqkv"].shape), tuple(shards["layers"][0]["wdown"].shape)))
    dist.destroy_process_group()


def test_tp_train_two_ranks_loss_and_grads_match_unsharded_hf():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs, s_qkv, s_down in res:
        assert s_qkv == (3 * 32, 64) and s_down == (64, 48)              # 2 of 4 heads x 16, half of the 96-wide MLP
        for k, v in errs.items():
            assert v < 2e-5, (rank, k, v)
