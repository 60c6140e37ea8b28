"""CPU: randomised-interleaving check of the tensor-parallel exchange protocol (visionllm_b200/tp.py + csrc/peer.cu).

The REAL call order of `TPLlamaForCausalLM.phases` is recorded through a comm double, then W ranks replay it as
asynchronous programs under thousands of random schedules with the protocol's semantics:
  * a push / scatter writes the sender's region of every peer's buffer, tagged with the sender's epoch, then bumps the
    peer's arrival counter; waits are enabled by `count >= epoch * arrivals` exactly like `spin_until`;
  * a consumer kernel reads for a DURATION (begin ... end, other ranks run in between), as GEMMs and the reduce kernel do;
  * safety = no region is overwritten while a reader of an older epoch is active on it, and every read sees exactly the
    epoch it waited for from every source; liveness = no schedule deadlocks.
The checker has teeth: dropping the counter barrier that opens a forward is caught as a hazard."""
import random
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F


class RecordingComm:
    """Records the comm-call trace of one rank; returns correctly shaped dummies."""

    def __init__(self, rows, hidden, world):
        self.rank, self.world, self.M, self.H, self.R = 0, world, rows, hidden, rows // world
        self.trace = []

    def barrier(self): self.trace.append("barrier")
    def barrier_wait(self): self.trace.append("barrier_wait")
    def norm_push(self, x, w, eps): self.trace.append("push")
    def gathered(self):
        self.trace.append("gather")
        return torch.zeros(self.M, self.H)
    def oproj_scatter(self, ctx, w): self.trace.append("scatter")
    def reduce_norm(self, x, w, eps):
        self.trace.append("reduce")
        return torch.zeros_like(x)


@pytest.fixture()
def recorded_trace(monkeypatch):
    from transformers import LlamaConfig
    import visionllm_b200.ops as ops
    from visionllm_b200 import tp
    monkeypatch.setattr(ops, "linear", lambda x, w, bias=None, act=None, residual=None, out=None, **kw:
                        (out if out is not None else torch.zeros(*x.shape[:-1], w.shape[0] // (2 if act == "swiglu" else 1))))
    monkeypatch.setattr(ops, "rope_", lambda *a, **k: None)
    monkeypatch.setattr(ops, "attention", lambda q, k, v, **kw: torch.zeros(q.shape[0], q.shape[1], q.shape[2] * q.shape[3]))
    cfg = LlamaConfig(hidden_size=32, intermediate_size=48, num_hidden_layers=3, num_attention_heads=4,
                      num_key_value_heads=4, vocab_size=20, rms_norm_eps=1e-5)
    world, B, T = 4, 2, 8
    comm = RecordingComm(B * T, 32, world)
    sd = {"embed": torch.zeros(20, 32), "final_norm": torch.ones(32), "lm_head": torch.zeros(20, 32),
          "layers": [{"wqkv": torch.zeros(3 * 8, 32), "wo": torch.zeros(32, 8), "w_gate_up": torch.zeros(96, 32),
                      "w_down": torch.zeros(32, 48), "ln1": torch.ones(32), "ln2": torch.ones(32)} for _ in range(3)]}
    m = tp.TPLlamaForCausalLM(cfg, comm)
    m.shards = sd
    m(inputs_embeds=torch.zeros(B, T, 32))
    return comm.trace


def expand(trace, n_forwards, with_barrier=True):
    """Comm calls -> atomic steps; consumers read for a duration: from their wait until the rank's next comm call."""
    prog = []
    for _ in range(n_forwards):
        open_read = None
        for op in trace:
            if open_read is not None and op not in ("barrier_wait",):
                prog.append(("end_read", open_read))
                open_read = None
            if op in ("barrier", "barrier_wait"):
                if with_barrier:
                    prog.append((op,))
            elif op in ("push", "scatter"):
                prog.append((op,))
            elif op == "gather":
                prog.append(("begin_read", "gather"))
                open_read = "gather"
            elif op == "reduce":
                prog.append(("begin_read", "recv"))
                prog.append(("end_read", "recv"))          # the reduce kernel itself is the whole read
        if open_read is not None:                          # the final gather is cloned before the forward returns
            prog.append(("end_read", open_read))
    return prog


def simulate(prog, W, seed, arrivals=3):
    rng = random.Random(seed)
    pc = [0] * W
    epoch = {k: [0] * W for k in ("gather", "recv", "barrier")}
    count = {k: [0] * W for k in ("gather", "recv", "barrier")}
    tag = {k: [[0] * W for _ in range(W)] for k in ("gather", "recv")}         # tag[buf][dst][src] = epoch written
    reading = {k: [None] * W for k in ("gather", "recv")}                       # epoch a rank is currently reading
    n = len(prog)

    def enabled(r):
        step = prog[pc[r]]
        if step[0] == "barrier_wait":
            return count["barrier"][r] >= epoch["barrier"][r] * W
        if step[0] == "begin_read":
            b = step[1]
            return count[b][r] >= epoch[b][r] * W * arrivals
        return True

    while any(p < n for p in pc):
        ready = [r for r in range(W) if pc[r] < n and enabled(r)]
        if not ready:
            return "deadlock"
        r = rng.choice(ready)
        step = prog[pc[r]]
        pc[r] += 1
        if step[0] == "barrier":
            epoch["barrier"][r] += 1
            for j in range(W):
                count["barrier"][j] += 1
        elif step[0] in ("push", "scatter"):
            b = "gather" if step[0] == "push" else "recv"
            epoch[b][r] += 1
            for j in range(W):
                if reading[b][j] is not None and reading[b][j] != epoch[b][r]:
                    return f"hazard: rank {r} overwrites {b}[{j}] epoch {reading[b][j]} under a reader"
                tag[b][j][r] = epoch[b][r]
                count[b][j] += arrivals
        elif step[0] == "begin_read":
            b = step[1]
            if any(t != epoch[b][r] for t in tag[b][r]):
                return f"stale: rank {r} reads {b} tags {tag[b][r]} at epoch {epoch[b][r]}"
            reading[b][r] = epoch[b][r]
        elif step[0] == "end_read":
            b = step[1]
            if any(t != reading[b][r] for t in tag[b][r]):
                return f"hazard: {b}[{r}] changed during the read of epoch {reading[b][r]}"
            reading[b][r] = None
    return "ok"


def test_recorded_trace_is_the_documented_phase_order(recorded_trace):
    L = 3
    assert recorded_trace == (["barrier", "barrier_wait", "push"] + ["gather", "scatter", "reduce", "push"] * L + ["gather"])


@pytest.mark.parametrize("W", [2, 4, 8])
def test_random_interleavings_are_safe_and_live(recorded_trace, W):
    prog = expand(recorded_trace, n_forwards=3)
    for seed in range(400):
        assert simulate(prog, W, seed) == "ok", (W, seed)


def test_checker_catches_a_missing_forward_barrier(recorded_trace):
    prog = expand(recorded_trace, n_forwards=3, with_barrier=False)
    verdicts = {simulate(prog, 4, seed) for seed in range(400)}
    assert any(v.startswith("hazard") or v.startswith("stale") for v in verdicts), verdicts
