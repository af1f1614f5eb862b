"""Drop-in for `prototype.model`: the same string-keyed factory (prototype/model/__init__.py:15-21).

    from declip_b200.model import model_entry
    model = model_entry(config.model)        # config.model.type in {'clip_vitb32', ...}

To run the reference solvers unchanged, alias the package before importing them:
    import sys, declip_b200.model as m; sys.modules['prototype.model'] = m        
"""
from .clip import CLIP, clip_res50, clip_vitb32  # noqa: F401
from .declip import DECLIP, declip_res50, declip_vitb32  # noqa: F401
from .defilip import DEFILIP, defilip_vitb32  # noqa: F401
from .filip import FILIP, filip_res50, filip_vitb32  # noqa: F401
from .slip import SLIP, slip_vitb32  # noqa: F401

_NOT_BUILT = ('slip_res50',)


def model_entry(config):
    name = config['type']
    if name in _NOT_BUILT:
        raise NotImplementedError("declip_b200: model type %r cannot run in the reference either (ModifiedResNet.forward "
                                  "has no return_feature argument, modified_resnet.py:192)" % name)
    if name not in globals():
        raise KeyError("unknown model type %r" % name)
    return globals()[name](**config['kwargs'])
