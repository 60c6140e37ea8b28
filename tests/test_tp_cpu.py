"""CPU, world_size 2 over gloo: host logic of the tensor-parallel Llama split (visionllm_b200/tp.py) -- head / row
ownership, packed-QKV and o_proj shard slicing, the reduce-scatter + all-gather exchange order, logits row ranges.

The CUDA kernels are replaced IN THIS TEST ONLY by fp32 torch (`ops.*`) and by a gloo-backed double of `PeerComm`
with the same five exchange methods; the expected values come from the installed HF `LlamaForCausalLM` (the third-party
module the reference instantiates, modeling_visionllmv2.py:143) run unsharded in fp32 on the same weights."""
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


class GlooComm:
    """Test double of tp.PeerComm: same methods, collectives over torch.distributed (gloo), fp32 math."""

    def __init__(self, rank, world, rows_total, hidden):
        self.rank, self.world, self.M, self.H = rank, world, rows_total, hidden
        self.R = rows_total // world
        self._gather = None
        self._partial = None

    def barrier(self):
        dist.barrier()

    def barrier_wait(self):
        pass

    @staticmethod
    def _rms(x, w, eps):
        return w.float() * (x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + eps))

    def norm_push(self, x_local, weight, eps):
        parts = [torch.zeros(self.R, self.H) for _ in range(self.world)]
        dist.all_gather(parts, self._rms(x_local, weight, eps).contiguous())
        self._gather = torch.cat(parts, 0)

    def gathered(self):
        return self._gather

    def oproj_scatter(self, ctx, w_shard):
        self._partial = F.linear(ctx.float(), w_shard.float())

    def reduce_norm(self, x_local, weight, eps):
        outs = torch.zeros(self.R, self.H)
        dist.reduce_scatter(outs, list(self._partial.split(self.R, 0)))
        x_local += outs
        return self._rms(x_local, weight, eps)


def _patch_ops():
    import visionllm_b200.ops as ops

    def linear(x, w, bias=None, act=None, colscale=None, residual=None, out_dtype=None, out=None):
        y = F.linear(x.float(), w.float(), None if bias is None else bias.float())
        if act == "swiglu":
            y = F.silu(y[..., 0::2]) * y[..., 1::2]
        elif act is not None:
            raise AssertionError(act)
        if residual is not None:
            y = y + residual.float().reshape(y.shape)
        if out is not None:
            out.copy_(y)
            return out
        return y

    def rope_(x, cos, sin, heads, D):
        v = x[:, :heads * D].reshape(x.shape[0], heads, D).float()
        rot = torch.cat((-v[..., D // 2:], v[..., :D // 2]), -1)
        x[:, :heads * D] = (v * cos[:, None, :].float() + rot * sin[:, None, :].float()).reshape(x.shape[0], heads * D)

    def attention(q, k, v, causal=False, scale=None, seqlens=None, **kw):
        B, T, Hq, D = q.shape
        rep = Hq // k.shape[2]
        kk, vv = k.repeat_interleave(rep, 2), v.repeat_interleave(rep, 2)
        s = (q.float().permute(0, 2, 1, 3) @ kk.float().permute(0, 2, 3, 1)) * D ** -0.5
        if causal:
            s = s.masked_fill(torch.ones(T, T).triu(1).bool(), float("-inf"))
        if seqlens is not None:
            s = s.masked_fill(torch.arange(T)[None, None, None, :] >= seqlens[:, None, None, None], float("-inf"))
        return (torch.softmax(s, -1) @ vv.float().permute(0, 2, 1, 3)).permute(0, 2, 1, 3).reshape(B, T, Hq * D)

    ops.linear, ops.rope_, ops.attention = linear, rope_, attention


def _config():
    from transformers import LlamaConfig
    return LlamaConfig(hidden_size=64, intermediate_size=96, num_hidden_layers=2, num_attention_heads=4,
                       num_key_value_heads=2, vocab_size=50, rms_norm_eps=1e-5, attn_implementation="eager")


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from transformers import LlamaForCausalLM
    from visionllm_b200 import tp
    _patch_ops()
    cfg = _config()
    torch.manual_seed(0)
    ref = LlamaForCausalLM(cfg).float().eval()
    B, T, H = 2, 8, cfg.hidden_size
    emb = torch.randn(B, T, H, generator=torch.Generator().manual_seed(1))
    am = torch.ones(B, T, dtype=torch.int64)
    am[1, 6:] = 0                                            # one right-padded sequence
    with torch.no_grad():
        want = ref(inputs_embeds=emb, attention_mask=am, output_hidden_states=True)
    comm = GlooComm(rank, world, B * T, H)
    model = tp.TPLlamaForCausalLM.from_full_state_dict(cfg, comm, ref.state_dict(), dtype=torch.float32)
    out = model(inputs_embeds=emb, attention_mask=am, output_hidden_states=True, use_cache=False)   # HF-style call
    assert out.hidden_states[-1] is out.last_hidden_state and out.logits is None
    assert model.dtype == torch.float32 and model.get_input_embeddings()(torch.tensor([[1, 2]])).shape == (1, 2, H)
    lo, hi = out.row_range
    valid = am.bool().reshape(-1)                            # padded query rows are don't-care in both
    err_h = (out.last_hidden_state.reshape(B * T, H) - want.hidden_states[-1].reshape(B * T, H))[valid].abs().max()
    err_l = (out.logits_local - want.logits.reshape(B * T, -1)[lo:hi])[valid[lo:hi]].abs().max()
    # shard bookkeeping
    ly = model.shards["layers"][0]
    D = H // cfg.num_attention_heads
    q.put((rank, (lo, hi), float(err_h), float(err_l), tuple(ly["wqkv"].shape), tuple(ly["wo"].shape),
           bool(torch.equal(ly["wo"], ref.state_dict()["model.layers.0.self_attn.o_proj.weight"]
                            [:, rank * 2 * D:(rank + 1) * 2 * D]))))
    dist.destroy_process_group()


def test_tp_llama_two_ranks_matches_unsharded_hf():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [(0, 8), (8, 16)]          # rows tile the flattened [B*T] axis exactly once
    for rank, _, err_h, err_l, wqkv, wo, wo_ok in res:
        assert err_h < 2e-5 and err_l < 2e-5, (rank, err_h, err_l)
        assert wqkv == ((2 + 1 + 1) * 16, 64) and wo == (64, 32) and wo_ok


def test_shard_rejects_indivisible_heads():
    import pytest
    from visionllm_b200 import tp
    cfg = _config()
    with pytest.raises(ValueError):
        tp.shard_llama_state_dict({}, cfg, 0, 3)


def test_exchange_layout():
    from visionllm_b200 import tp
    assert tp.exchange_bytes(4096, 4096, 8) == 4096 + 2 * 4096 * 4096 * 2


def test_internlm2_shards_equal_llama_shards_of_the_unfused_weights():
    """The reference's fused InternLM2 `wqkv` ((G q heads, k, v) per KV head, internlm2/modeling_internlm2.py:337-349)
    un-fused into HF-Llama q/k/v projections must shard to exactly the same per-rank tensors: GQA groups stay
    rank-local and the rank's query heads are the contiguous block `wo` is ordered by."""
    from types import SimpleNamespace
    from visionllm_b200 import tp
    nq, nkv, D, H, I, L, V = 8, 4, 4, 32, 24, 2, 11
    G = nq // nkv
    cfg = SimpleNamespace(num_attention_heads=nq, num_key_value_heads=nkv, hidden_size=H, num_hidden_layers=L)
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    il, ll = {"model.tok_embeddings.weight": r(V, H), "model.norm.weight": r(H), "output.weight": r(V, H)}, {}
    ll.update({"model.embed_tokens.weight": il["model.tok_embeddings.weight"], "model.norm.weight": il["model.norm.weight"],
               "lm_head.weight": il["output.weight"]})
    for i in range(L):
        q, k, v = r(nq * D, H), r(nkv * D, H), r(nkv * D, H)
        fused = torch.cat([torch.cat([q.view(nkv, G * D, H)[j], k.view(nkv, D, H)[j], v.view(nkv, D, H)[j]], 0)
                           for j in range(nkv)], 0)
        p, pl = f"model.layers.{i}.", f"model.layers.{i}."
        il.update({p + "attention.wqkv.weight": fused, p + "attention.wo.weight": r(H, nq * D),
                   p + "feed_forward.w1.weight": r(I, H), p + "feed_forward.w3.weight": r(I, H),
                   p + "feed_forward.w2.weight": r(H, I), p + "attention_norm.weight": r(H), p + "ffn_norm.weight": r(H)})
        ll.update({pl + "self_attn.q_proj.weight": q, pl + "self_attn.k_proj.weight": k, pl + "self_attn.v_proj.weight": v,
                   pl + "self_attn.o_proj.weight": il[p + "attention.wo.weight"],
                   pl + "mlp.gate_proj.weight": il[p + "feed_forward.w1.weight"],
                   pl + "mlp.up_proj.weight": il[p + "feed_forward.w3.weight"],
                   pl + "mlp.down_proj.weight": il[p + "feed_forward.w2.weight"],
                   pl + "input_layernorm.weight": il[p + "attention_norm.weight"],
                   pl + "post_attention_layernorm.weight": il[p + "ffn_norm.weight"]})
    for world in (1, 2, 4):
        for rank in range(world):
            a = tp.shard_internlm2_state_dict(il, cfg, rank, world)
            b = tp.shard_llama_state_dict(ll, cfg, rank, world)
            for key in ("embed", "final_norm", "lm_head"):
                assert torch.equal(a[key], b[key])
            for la, lb in zip(a["layers"], b["layers"]):
                assert la.keys() == lb.keys()
                for key in la:
                    assert torch.equal(la[key], lb[key]), (world, rank, key)
    import pytest
    with pytest.raises(ValueError):
        tp.shard_internlm2_state_dict(il, cfg, 0, 8)            # 4 KV heads over 8 ranks
