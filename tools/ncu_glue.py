"""The glue kernels of the second half of round 2 at their GDINO / UniPose shapes, one launch each between
cudaProfilerStart/Stop, for one `ncu --set full` capture:
    1. gather_rows_flat_kernel   Swin stage-0 window reverse: [8 * 65856 window slots, 96] -> [8 * 65536, 96]
    2. sine_embed_kernel         neck level 0: 8 x 128 x 128 pixels -> bf16 [.., 256] + level embedding, into a [8, 21760, 256] slab
    3. sine_embed_kernel         decoder proposals: 8 x 100 boxes -> bf16 [800, 512]
    4. upsample_add_nhwc_kernel  mask FPN: top [8, 128, 128, 256] (batch pitch of the flattened encoder output) -> padded [8, 258, 258, 256]
    5. gn_stats / gn_apply       GroupNorm(32, 256) + ReLU on the valid 256 x 256 corner of that padded grid
    6. mask_tiles_kernel         UniPose keypoint mask [32, 3450, 3450] -> live-tile lists
    7. flash_fwd_kernel<32, 64>  the keypoint self-attention walking those lists (B=4, H=8, T=3450, D=32)

  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_glue python tools/ncu_glue.py
  python tools/ncu_summary.py full gpurun_out/r2_glue.ncu-rep profiles/r2_glue_kernels_ncu.json
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from visionllm_b200 import ops
    from visionllm_b200.gdino_model import GroundingDinoSinePositionEmbedding
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    r = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()  # noqa: E731
    src = r(8 * 65856, 96)
    idx = torch.randint(0, 8 * 65856, (8 * 65536,), device=dev, generator=g)
    pe = GroundingDinoSinePositionEmbedding(128, 20, normalize=True)
    m0 = torch.ones(8, 128, 128, dtype=torch.bool, device=dev)
    y_e, x_e = (t.contiguous() for t in pe.embeds(m0))
    lvl = r(256)
    flat = torch.empty(8, 21760, 256, dtype=torch.bfloat16, device=dev)
    ref_in = torch.rand(8, 100, 4, 4, device=dev, generator=g)
    p = ref_in[:, :, 0, :]
    d = torch.arange(128, dtype=torch.float32, device=dev)
    dim_t = 10000 ** (2 * torch.div(d, 2, rounding_mode="floor") / 128)
    enc = r(8, 21760, 256)
    top = enc[:, :128 * 128].reshape(8, 128, 128, 256)
    lat = r(8, 256, 256, 256)
    gam, bet = r(256), r(256)
    grid = r(8, 258, 258, 256)[:, :256, :256]
    T = 3450
    ii = torch.arange(T, device=dev)
    allow = ((ii[:, None] // 69) == (ii[None, :] // 69))[None].repeat(32, 1, 1)
    q, k, v = (r(4, T, 8, 32) for _ in range(3))

    def targets(tiles):
        ops.gather_rows(src, idx)
        ops.sine_embed([y_e, x_e], 1, pe.dim_t(dev), 8 * 128 * 128, out=flat[:, :128 * 128], add_row=lvl)
        ops.sine_embed([p[:, :, c] for c in (1, 0, 2, 3)], p.stride(1), dim_t, 800, pre_scale=6.283185307179586, out_dtype=torch.bfloat16)
        ops.upsample_add_nhwc(top, lat, pad=1)
        ops.groupnorm_nhwc(grid, gam, bet, 32, 1e-5, relu=True)
        t = ops.attention_mask_tiles(allow)
        ops.attention(q, k, v, attn_mask=t)

    targets(None)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    targets(None)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("done")


if __name__ == "__main__":
    main()
