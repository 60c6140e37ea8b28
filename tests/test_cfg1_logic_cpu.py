"""CPU: BASELINE cfg 1 end to end through `B200VisionLLMv2Model.forward` with the CUDA kernels replaced -- in this test
only -- by the fp32 torch stand-ins of oracle/torch_kernels.py, against the golden produced by the REFERENCE's own
modules (tests/golden/gen_golden_cfg1.py).  fp32 against fp32: the composite's host logic ([EMB] injection, image-token
scatter, text_query gather, Swin/neck/encoder/decoder wiring, two-stage top-k) must reproduce the reference to
rounding noise, indices exactly.  The GPU kernels themselves are checked by tests/test_cfg1_e2e_gpu.py."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
sys.path.insert(0, os.path.dirname(__file__))


def build_cfg1_fp32_cpu(g):
    import cfg1_common as C
    return C.build_b200_model(g, device="cpu", dtype=torch.float32)


def test_cfg1_composite_host_logic_matches_reference_modules(golden_dir):
    import cfg1_common as C
    from oracle import torch_kernels as TK
    g = np.load(os.path.join(golden_dir, "cfg1_e2e.npz"))
    model = build_cfg1_fp32_cpu(g)
    ids, image, aug = C.inputs()
    with TK.patched():
        out = model(input_ids=ids, attention_mask=torch.ones_like(ids), images=image, images_aug=[aug[0]],
                    img_metas=[{"task": "det"}])

    def rel(a, name):
        b = torch.from_numpy(g[name + "_f32"])
        m = torch.isfinite(b)
        return ((a.float().reshape(b.shape)[m] - b[m]).norm() / b[m].norm()).item()

    assert torch.equal(out.input_ids, torch.from_numpy(g["new_input_ids"]))
    assert rel(out.logits, "llm_logits") < 1e-4 and rel(out.last_hidden_state, "llm_hidden") < 1e-4
    go = out.gdino_outputs
    assert torch.equal(go.model_outputs.topk_proposals, torch.from_numpy(g["topk"]))     # two-stage selection, exact
    assert rel(go.model_outputs.enc_outputs_class.float().max(-1)[0], "enc_class_max") < 1e-4
    assert rel(go.logits, "gd_logits") < 1e-4 and rel(go.pred_boxes, "gd_boxes") < 1e-4
    assert rel(go.pred_masks.reshape(1, -1)[:, ::int(g["mask_sub"])], "gd_masks") < 1e-4
    # detection post-processing: the golden's indices (eval_det.py:18-56 on the reference outputs), exact
    from visionllm_b200 import gdino_heads as H
    res, topk_indexes, box_idx = H.post_process_det_gdino(go.logits, go.pred_boxes, [(224, 224)], C.N_CLS, topk=100)
    assert torch.equal(topk_indexes, torch.from_numpy(g["det_topk_indexes"]))
    assert torch.equal(box_idx, torch.from_numpy(g["det_box_idx"])) and torch.equal(res[0]["labels"], torch.from_numpy(g["det_labels"])[0])
