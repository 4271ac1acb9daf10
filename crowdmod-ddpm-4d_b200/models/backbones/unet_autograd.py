"""Autograd bridge for the native UNet (training forward/backward).

Round-1 status: the native backward kernels (dgrad / wgrad on tcgen05, GroupNorm / SiLU /
attention backward) are not implemented yet, so training through the native path fails
loudly instead of silently falling back to eager PyTorch.
"""


def unet_train_forward(module, future, t, past):
    raise NotImplementedError(
        "crowdmod-ddpm-4d_b200: the native training backward (SURVEY.md §8 a7) is not built "
        "yet; run the UNet under torch.no_grad()/inference_mode() for sampling")
