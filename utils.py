"""Drop-in for the two helpers of the reference's ``utils.py`` that its trainer uses (:23-31).
The rest of that file (bicubic resampling, init functions, TVLoss; needs lpips/matplotlib) is not on
the training hot path and is not reproduced here."""


def freeze(model):
    """requires_grad=False on every parameter + eval() (no BatchNorm/Dropout exists, so the mode has
    no numeric effect; a frozen T_net means the F-sub forward builds no graph)."""
    for p in model.parameters():
        p.requires_grad_(False)
    model.eval()


def unfreeze(model):
    for p in model.parameters():
        p.requires_grad_(True)
    model.train(True)
