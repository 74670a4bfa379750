def rasterization(*args, **kwargs):
    raise NotImplementedError("gsplat.rendering.rasterization is not part of the vicasplat_b200 hot path "
                              "(every shipped experiment sets decoder.use_gsplat: false)")
