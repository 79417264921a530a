"""adaprompt_b200 - B200-native (sm_100a) implementation of the AdaFace / SD-1.5 denoising hot path.

Host side mirrors the reference's module API (askerlee/adaprompt: ldm.modules.diffusionmodules.openaimodel,
ldm.modules.attention, ldm.models.diffusion.ddim, adaface.subj_basis_generator); all arithmetic runs in
hand-written CUDA behind the C ABI declared in include/adaface_b200.h.
"""
__version__ = "0.1.0"
