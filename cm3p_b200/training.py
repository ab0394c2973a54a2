"""Explicit-backward training path (filled in by the backward kernels)."""


def forward_with_grad(model, **kwargs):
    raise NotImplementedError(
        "cm3p_b200: the training (gradient) path is not built yet; call the model under torch.no_grad() "
        "for inference")
