"""Minimal hydra stand-in: `@hydra.main`, `hydra.utils`, `hydra.core.hydra_config.HydraConfig`.

`main(version_base=None, config_path, config_name)` loads `<config_path>/<config_name>.yaml`
relative to the decorated function's file, applies `a.b=value` command-line overrides (values
parsed as YAML), creates `outputs/<date>/<time>/` as the run's output dir and does not change
the working directory (what Hydra >= 1.2 does with version_base=None; the reference's config
uses paths relative to the repo root)."""
import datetime
import functools
import inspect
import os
import sys

import yaml
from omegaconf import DictConfig, OmegaConf

from . import utils  # noqa: F401
from .core import hydra_config
from . import core  # noqa: F401

__version__ = "0.0-ddope-shim"


def main(version_base=None, config_path=None, config_name=None):
    def decorator(fn):
        @functools.wraps(fn)
        def wrapper(cfg_passthrough=None):
            if cfg_passthrough is not None:
                return fn(cfg_passthrough)
            base = os.path.dirname(os.path.abspath(inspect.getsourcefile(fn)))
            cfg = DictConfig()
            if config_name is not None:
                name = config_name if config_name.endswith((".yaml", ".yml")) else config_name + ".yaml"
                cfg = OmegaConf.load(os.path.normpath(os.path.join(base, config_path or ".", name)))
            out_override = None
            for arg in sys.argv[1:]:
                if "=" not in arg:
                    continue
                key, val = arg.split("=", 1)
                key = key.lstrip("+")
                if key == "hydra.run.dir":
                    out_override = val
                    continue
                OmegaConf.update(cfg, key, yaml.safe_load(val))
            now = datetime.datetime.now()
            out_dir = out_override or os.path.join(os.getcwd(), "outputs", now.strftime("%Y-%m-%d"), now.strftime("%H-%M-%S"))
            os.makedirs(out_dir, exist_ok=True)
            hydra_config.HydraConfig._set(DictConfig({"runtime": {"output_dir": out_dir, "cwd": os.getcwd()}, "job": {"name": fn.__name__}}))
            with open(os.path.join(out_dir, "config.yaml"), "w") as f:
                f.write(OmegaConf.to_yaml(cfg))
            return fn(cfg)

        return wrapper

    return decorator
