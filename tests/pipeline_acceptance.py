"""Pipeline-level drop-in acceptance: the REAL `CogVideoXI2VDualInpaintAnyLPipeline.__call__` (PIPE:633-1083)
of the reference — tiny random VAE, no tokenizer / T5, a tiny transformer and a 1-layer branch — on two chained 49-frame
windows of 4 steps each (`prev_clip_weight = 0.5`, `replace_gt`, `mask_add`, dynamic CFG, DPM scheduler), with every
`noise_pred` the pipeline receives from `self.transformer(...)` recorded by a forward hook.

Test infrastructure.  The reference code is imported from wherever it is available:
  * /root/reference/diffusers/src              (the build container)
  * baseline/_ref                              (`pip install --target baseline/_ref` of the reference's diffusers fork; git-ignored,
                                                travels to the GPU box with the repo snapshot — DESIGN.md §8)

    python tests/pipeline_acceptance.py        # CPU fp32, reference forwards -> tests/golden/pipeline_tiny.pt

The GPU half (tests/test_gpu_pipeline.py) runs the same pipeline on cuda in bf16 twice — with the reference forwards and with
`videopainter_b200.install()` — and compares every recorded call."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden", "pipeline_tiny.pt")
_CANDIDATES = ["/root/reference/diffusers/src", os.path.join(ROOT, "baseline", "_ref")]


def reference_path():
    for p in _CANDIDATES:
        if os.path.isdir(os.path.join(p, "diffusers", "pipelines", "cogvideo")):
            return p
    return None


def load_reference():
    """Import the reference's pipeline, VAE and scheduler classes (None if the reference is not available here)."""
    p = reference_path()
    if p is None:
        return None
    if p not in sys.path:
        sys.path.insert(0, p)
    import transformers.utils as tu
    if not hasattr(tu, "FLAX_WEIGHTS_NAME"):                  # transformers 5 dropped it; pipeline_loading_utils.py:49 imports it
        tu.FLAX_WEIGHTS_NAME = "flax_model.msgpack"
    from diffusers import (AutoencoderKLCogVideoX, CogVideoXDPMScheduler, CogVideoXTransformer3DModel,  # type: ignore
                           CogvideoXBranchModel)
    from diffusers.pipelines.cogvideo.pipeline_cogvideox_inpainting_i2v_branch_anyl import (  # type: ignore
        CogVideoXI2VDualInpaintAnyLPipeline)
    return dict(pipe=CogVideoXI2VDualInpaintAnyLPipeline, vae=AutoencoderKLCogVideoX, sched=CogVideoXDPMScheduler,
                transformer=CogVideoXTransformer3DModel, branch=CogvideoXBranchModel)


def build_pipeline(ref, device="cpu", dtype=torch.float32, resample=False):
    from oracle import cogvideox_oracle as O
    cfg = O.tiny_config(id_pool_resample_learnable=resample)
    cfg_b = O.tiny_config(num_layers=1)
    kw = cfg.to_kwargs(); kw.pop("norm_eps")
    kwb = cfg_b.to_kwargs(); kwb.pop("norm_eps"); kwb.pop("id_pool_resample_learnable")
    tr = ref["transformer"](**kw)
    tr.load_state_dict(O.init_state_dict(cfg, 71), strict=True)
    br = ref["branch"](**kwb)
    br.load_state_dict(O.init_state_dict(cfg_b, 72, branch=True), strict=True)
    torch.manual_seed(7)
    vae = ref["vae"](in_channels=3, out_channels=3, down_block_types=("CogVideoXDownBlock3D",) * 4,
                     up_block_types=("CogVideoXUpBlock3D",) * 4, block_out_channels=(8, 8, 8, 8), latent_channels=16,
                     layers_per_block=1, norm_num_groups=2, temporal_compression_ratio=4, scaling_factor=0.7)
    sched = ref["sched"](snr_shift_scale=1.0, prediction_type="v_prediction", rescale_betas_zero_snr=True,
                         timestep_spacing="trailing", clip_sample=False)
    pipe = ref["pipe"](tokenizer=None, text_encoder=None, vae=vae, transformer=tr, scheduler=sched, branch=br)
    pipe = pipe.to(device=device, dtype=dtype)
    pipe.set_progress_bar_config(disable=True)
    return pipe


def make_inputs(n_frames=98, size=64):
    """98 synthetic RGB frames (two 49-frame windows at stride 49) and masks: a moving rectangle, frame 0 all black."""
    import numpy as np
    from PIL import Image
    g = np.random.RandomState(3)
    video, masks = [], []
    base = g.randint(0, 255, size=(size, size, 3)).astype(np.uint8)
    for f in range(n_frames):
        frame = np.roll(base, shift=f, axis=1).copy()
        frame[(f * 3) % size] = 255
        video.append(Image.fromarray(frame))
        m = np.zeros((size, size), dtype=np.uint8)
        if f > 0:
            y0, x0 = (f * 2) % (size // 2), (f * 3) % (size // 2)
            m[y0:y0 + size // 2, x0:x0 + size // 2] = 255
        masks.append(Image.fromarray(m).convert("RGB"))
    ge = torch.Generator().manual_seed(11)
    prompt = torch.randn(1, 16, 64, generator=ge)
    negative = torch.randn(1, 16, 64, generator=ge)
    return video, masks, prompt, negative


def run(pipe, resample=False):
    """Returns (list of every noise_pred the transformer handed to the pipeline, final latents)."""
    preds = []
    handle = pipe.transformer.register_forward_hook(lambda mod, args, out: preds.append(out[0].detach().float().cpu().clone()))
    video, masks, prompt, negative = make_inputs()
    dev, dt = pipe.transformer.device, pipe.transformer.dtype
    try:
        with torch.no_grad():
            out = pipe(image=video[0], prompt=None, prompt_embeds=prompt.to(dev, dt), negative_prompt_embeds=negative.to(dev, dt),
                       max_sequence_length=16, height=64, width=64, num_frames=49, num_inference_steps=4, guidance_scale=6,
                       use_dynamic_cfg=True, video=video, masks=masks, strength=1.0, replace_gt=True, mask_add=True, stride=49,
                       prev_clip_weight=0.5, id_pool_resample_learnable=resample, output_type="latent",
                       generator=torch.Generator().manual_seed(42))
    finally:
        handle.remove()
    lat = out.frames if hasattr(out, "frames") else out[0]
    return preds, lat.detach().float().cpu()


def main():
    ref = load_reference()
    if ref is None:
        raise SystemExit("the reference is not available here (neither /root/reference nor baseline/_ref)")
    torch.set_num_threads(os.cpu_count())
    rec = {}
    for resample in (False, True):
        pipe = build_pipeline(ref, "cpu", torch.float32, resample)
        preds, lat = run(pipe, resample)
        tag = "resample" if resample else "plain"
        rec[tag] = dict(noise_preds=preds, latents=lat)
        print(tag, len(preds), "transformer calls, latents", tuple(lat.shape), "absmax", float(lat.abs().max()))
    torch.save(rec, GOLDEN)
    print("->", GOLDEN, os.path.getsize(GOLDEN) // 1024, "KiB")


if __name__ == "__main__":
    main()
