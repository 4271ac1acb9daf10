"""YAML config loader — drop-in for reference utils/myparser.py:1-33, plus a schema adapter.

``getYamlConfig(config_yml_file, configList_yml_file)`` merges the two YAML files into one
attribute-accessible dict exactly like the reference (which uses EasyDict).  In addition,
``adapt_legacy_schema`` upgrades the legacy flat schema used by three of the five BASELINE
configs (``config/ATC_synthetic.yml``, ``ATC_medium.yml``, ``ETHUCY_ddpm.yml``: top-level
``DIFFUSION:`` / ``TRAIN:`` and ``MODEL.BASE_CH``) to the nested ``MODEL.DDPM.UNET`` schema the
current reference code reads (models/diffusion/ddpm.py:65-72), and fills the two keys newer
code needs but some configs lack (``CHECKPOINTS_TO_KEEP`` <- ``MODEL_SAMPLES`` in
config/HERMES-CR-120.yml:50; ``LAMBDA_GUIDANCE`` default 0.004 from config/HERMES-BO.yml:50).
"""
import os

import yaml


class YamlParser(dict):
    """dict with recursive attribute access (EasyDict semantics)."""

    def __init__(self, cfg_dict=None, config_file=None):
        super().__init__()
        cfg_dict = dict(cfg_dict or {})
        if config_file is not None:
            assert os.path.isfile(config_file)
            with open(config_file, 'r') as fo:
                cfg_dict.update(yaml.safe_load(fo.read()))
        for k, v in cfg_dict.items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, YamlParser):
            return YamlParser(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def update(self, other=None, **kw):
        for k, v in dict(other or {}, **kw).items():
            self[k] = v

    def merge_from_file(self, config_file):
        with open(config_file, 'r') as fo:
            self.update(yaml.safe_load(fo.read()))

    def merge_from_dict(self, config_dict):
        self.update(config_dict)


def get_config(config_file=None):
    return YamlParser(config_file=config_file)


_UNET_KEYS = ("CONDITION", "CONDITION_HANDLING", "NUM_RES_BLOCKS", "BASE_CH", "BASE_CH_MULT",
              "APPLY_ATTENTION", "DROPOUT_RATE", "TIME_EMB_MULT")
_DDPM_KEYS = ("SAMPLER", "TIMESTEPS", "SCALE", "SIGMA", "DDIM_DIVIDER", "GUIDANCE",
              "LAMBDA_GUIDANCE", "CHECKPOINTS_TO_KEEP")


def adapt_legacy_schema(cfg):
    """In-place upgrade legacy-flat -> nested schema; a no-op on already-nested configs."""
    model = cfg.get("MODEL")
    if model is None:
        return cfg
    if "DDPM" not in model and "BASE_CH" in model:
        diffusion = cfg.get("DIFFUSION", {})
        train = cfg.get("TRAIN", {})
        unet = {k: model[k] for k in _UNET_KEYS if k in model}
        solver = dict(train.get("SOLVER", {}))
        solver.setdefault("LR", train.get("INITIAL_LR", 5e-5))
        solver.setdefault("WEIGHT_DECAY", 0.0)
        solver.setdefault("BETAS", [0.9, 0.999])
        solver.setdefault("SCHEDULER", {"FACTOR": 0.5, "PATIENCE": 10, "MIN_LR": 1e-6})
        unet["TRAIN"] = {"EPOCHS": train.get("EPOCHS", 1), "SOLVER": solver}
        ddpm = {k: diffusion[k] for k in _DDPM_KEYS if k in diffusion}
        ddpm["UNET"] = unet
        model["DDPM"] = ddpm
        for k in ("NSAMPLES", "NSAMPLES4PLOTS"):
            if k in diffusion and k not in model:
                model[k] = diffusion[k]
        if "NAME" not in model:
            model["NAME"] = "{}_" + str(cfg.get("DATASET", {}).get("NAME", "DS")) + "_TE{}_PL{}_FL{}_CE{}_{}.pth"
        fs = cfg.get("DATA_FS", {})
        fs = dict(fs) if fs else {}
        fs.setdefault("SAVE_DIR", model.get("SAVE_DIR", "saved_models/"))
        fs.setdefault("OUTPUT_DIR", model.get("OUTPUT_DIR", "output"))
        if "PICKLE" in cfg:
            fs.setdefault("PICKLE_DIR", cfg.PICKLE.get("PICKLE_DIR", ""))
            fs.setdefault("USE_PICKLE", cfg.PICKLE.get("USE_PICKLE", False))
        cfg["DATA_FS"] = fs
    ddpm = cfg.MODEL.get("DDPM")
    if ddpm is not None:
        if "CHECKPOINTS_TO_KEEP" not in ddpm:
            ddpm["CHECKPOINTS_TO_KEEP"] = ddpm.get("MODEL_SAMPLES", 7)
        ddpm.setdefault("LAMBDA_GUIDANCE", 0.004)
        ddpm.setdefault("SIGMA", 0.001)
        ddpm.setdefault("SAMPLER", "DDPM")
        ddpm.setdefault("GUIDANCE", "None")
        ddpm.setdefault("DDIM_DIVIDER", 2)
        if str(ddpm.GUIDANCE).lower() == "sparsity":
            ddpm["GUIDANCE"] = "Sparsity"
    return cfg


def getYamlConfig(config_yml_file, configList_yml_file=None, adapt=True):
    cfg = get_config()
    cfg.merge_from_file(config_file=config_yml_file)
    if configList_yml_file is not None:
        cfg.merge_from_file(config_file=configList_yml_file)
    return adapt_legacy_schema(cfg) if adapt else cfg
