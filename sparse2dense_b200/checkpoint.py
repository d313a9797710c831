"""Checkpoint interop with the reference format (det3d/torchie/trainer/checkpoint.py:146-240): a checkpoint is
``{"meta": dict, "state_dict": cpu tensors, ["optimizer": ...]}``; loading accepts that dict or a bare (Ordered)Dict of
tensors, strips a leading ``module.`` (DistributedDataParallel), unwraps ``model.module`` and is non-strict by default,
reporting what did not match.  Parameter names and the spconv weight layout ``[kD,kH,kW,Cin,Cout]`` of this package are
the reference's, so reference ``epoch_N.pth`` files load."""
import logging
import os
import time
from collections import OrderedDict

import torch


def load_state_dict(module, state_dict, strict=False, logger=None):
    own = module.state_dict()
    unexpected, mismatch = [], []
    for name, param in state_dict.items():
        if name not in own:
            unexpected.append(name)
            continue
        if tuple(param.shape) != tuple(own[name].shape):
            mismatch.append("{}: checkpoint {} vs model {}".format(name, tuple(param.shape), tuple(own[name].shape)))
            continue
        own[name].copy_(param.data if isinstance(param, torch.nn.Parameter) else param)
    missing = sorted(set(own.keys()) - set(state_dict.keys()))
    msgs = []
    if unexpected:
        msgs.append("unexpected key in source state_dict: {}".format(", ".join(unexpected)))
    if missing:
        msgs.append("missing keys in source state_dict: {}".format(", ".join(missing)))
    if mismatch:
        msgs.append("size mismatch: {}".format("; ".join(mismatch)))
    if msgs:
        text = "The model and loaded state dict do not match exactly\n" + "\n".join(msgs)
        if strict:
            raise RuntimeError(text)
        (logger or logging.getLogger(__name__)).warning(text)
    return dict(unexpected=unexpected, missing=missing, mismatch=mismatch)


def load_checkpoint(model, filename, map_location=None, strict=False, logger=None):
    if not os.path.isfile(filename):
        raise IOError("{} is not a checkpoint file".format(filename))
    checkpoint = torch.load(filename, map_location=map_location)
    if isinstance(checkpoint, OrderedDict):
        state_dict = checkpoint
    elif isinstance(checkpoint, dict) and "state_dict" in checkpoint:
        state_dict = checkpoint["state_dict"]
    else:
        raise RuntimeError("No state_dict found in checkpoint file {}".format(filename))
    if len(state_dict) and list(state_dict.keys())[0].startswith("module."):
        state_dict = {k[7:]: v for k, v in state_dict.items()}
    load_state_dict(model.module if hasattr(model, "module") else model, state_dict, strict, logger)
    return checkpoint


def save_checkpoint(model, filename, optimizer=None, meta=None):
    if meta is None:
        meta = {}
    elif not isinstance(meta, dict):
        raise TypeError("meta must be a dict or None, but got {}".format(type(meta)))
    meta = dict(meta)
    meta.setdefault("time", time.asctime())
    d = os.path.dirname(filename)
    if d:
        os.makedirs(d, exist_ok=True)
    if hasattr(model, "module"):
        model = model.module
    checkpoint = {"meta": meta, "state_dict": OrderedDict((k, v.cpu()) for k, v in model.state_dict().items())}
    if optimizer is not None:
        checkpoint["optimizer"] = optimizer.state_dict()
    torch.save(checkpoint, filename)
