"""Config surface of the reference (det3d/torchie/utils/config.py:51-161): python-file configs loaded as modules and
exposed through a dict with attribute access.  The reference builds ``ConfigDict`` on the third-party ``addict``
package; this one is self-contained (same behaviour where the configs and builders rely on it: nested dicts become
ConfigDicts, a missing key raises AttributeError / KeyError, ``.get`` works)."""
import importlib.util
import os
import sys


class ConfigDict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setitem__(self, key, value):
        super().__setitem__(key, self._wrap(value))

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError("'{}' object has no attribute '{}'".format(type(self).__name__, name))

    def __setattr__(self, name, value):
        self[name] = value

    def __deepcopy__(self, memo):
        import copy
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


class Config(object):
    @staticmethod
    def fromfile(filename):
        filename = os.path.abspath(os.path.expanduser(filename))
        if not os.path.isfile(filename):
            raise FileNotFoundError('file "{}" does not exist'.format(filename))
        if not filename.endswith(".py"):
            raise IOError("Only py type is supported by this build (the Waymo configs are python files)")
        module_name = os.path.basename(filename)[:-3]
        if "." in module_name:
            raise ValueError("Dots are not allowed in config file path.")
        sys.path.insert(0, os.path.dirname(filename))
        try:
            spec = importlib.util.spec_from_file_location("_s2d_cfg_" + module_name, filename)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        finally:
            sys.path.pop(0)
        cfg_dict = {k: v for k, v in mod.__dict__.items() if not k.startswith("__") and not isinstance(v, type(os))}
        return Config(cfg_dict, filename=filename)

    def __init__(self, cfg_dict=None, filename=None):
        if cfg_dict is None:
            cfg_dict = dict()
        elif not isinstance(cfg_dict, dict):
            raise TypeError("cfg_dict must be a dict, but got {}".format(type(cfg_dict)))
        object.__setattr__(self, "_cfg_dict", ConfigDict(cfg_dict))
        object.__setattr__(self, "_filename", filename)
        text = ""
        if filename:
            with open(filename, "r") as f:
                text = f.read()
        object.__setattr__(self, "_text", text)

    @property
    def filename(self):
        return self._filename

    @property
    def text(self):
        return self._text

    def __repr__(self):
        return "Config (path: {}): {}".format(self.filename, self._cfg_dict.__repr__())

    def __len__(self):
        return len(self._cfg_dict)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __setattr__(self, name, value):
        self._cfg_dict[name] = value

    def __setitem__(self, name, value):
        self._cfg_dict[name] = value

    def __iter__(self):
        return iter(self._cfg_dict)

    def __contains__(self, name):
        return name in self._cfg_dict

    def get(self, name, default=None):
        return self._cfg_dict.get(name, default)


def get_downsample_factor(model_config):
    """det3d/utils/config_tool.py:39-53: prod(ds_layer_strides) / us_layer_strides[-1] * backbone.ds_factor; a
    two-stage config is resolved through its ``first_stage_cfg``."""
    if "neck" not in model_config:
        model_config = model_config["first_stage_cfg"]
    neck = model_config["neck"]
    factor = 1.0
    for s in neck.get("ds_layer_strides", [1]):
        factor *= s
    ups = neck.get("us_layer_strides", [])
    if len(ups) > 0:
        factor /= ups[-1]
    factor = int(factor * model_config["backbone"]["ds_factor"])
    assert factor > 0
    return factor
