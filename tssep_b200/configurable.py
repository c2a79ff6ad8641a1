"""Minimal stand-in for ``padertorch.Configurable`` (padertorch is not a
dependency of this package).

The reference instantiates every hot-path object from a config dict whose
``factory`` key names the class (tssep/exp/init_cfg_common.yaml:7-83,
README.md:98-99 "you can replace the factorys in the config with your own
classes").  This module provides the five entry points the reference relies
on -- ``get_config``, ``from_config``, ``new``, ``finalize_dogmatic_config`` and
``import_class`` -- with the dict layout its doctests pin
(tssep/train/net.py:66-107, tssep/train/model.py:74-115): ``'factory'`` first,
then every constructor argument with its default, nested factories expanded.

``remap_factories`` rewrites the reference's factory strings to this package's
drop-in classes so the shipped YAML files load unchanged.
"""
from __future__ import annotations

import collections.abc
import copy
import importlib
import inspect
from typing import Any, Dict

# reference factory -> drop-in class of this package
FACTORY_ALIASES = {
    "tssep.train.model.Model": "tssep_b200.model.Model",
    "tssep.train.feature_extractor.ConcaternatedSTFTFeatures": "tssep_b200.feature_extractor.ConcaternatedSTFTFeatures",
    "tssep.train.feature_extractor.Log1pMaxNormAbsSTFT": "tssep_b200.feature_extractor.Log1pMaxNormAbsSTFT",
    "tssep.train.feature_extractor.TorchMFCC": "tssep_b200.feature_extractor_torchaudio.TorchMFCC",
    "tssep.train.enhancer.ClassicBF_np": "tssep_b200.enhancer.ClassicBF_np",
    "tssep.train.enhancer_distortion_mask.SumCrossTalker": "tssep_b200.enhancer_distortion_mask.SumCrossTalker",
    "tssep.train.enhancer_distortion_mask.OneMinus": "tssep_b200.enhancer_distortion_mask.OneMinus",
    "tssep.train.enhancer.WPE": "tssep_b200.enhancer.WPE",
    "tssep.train.enhancer.ChannelWiseWPE": "tssep_b200.enhancer.ChannelWiseWPE",
    "tssep.train.feature_extractor.Log1pAbsIPDSTFT": "tssep_b200.feature_extractor.Log1pAbsIPDSTFT",
    "tssep.train.feature_extractor.Log1pMaxNormAbsIPDSTFT": "tssep_b200.feature_extractor.Log1pMaxNormAbsIPDSTFT",
    "tssep.train.feature_extractor.MVNLog1pAbsSTFT": "tssep_b200.feature_extractor.MVNLog1pAbsSTFT",
    "tssep.train.feature_extractor.NoFeatureSTFT": "tssep_b200.feature_extractor.NoFeatureSTFT",
    "tssep.train.feature_extractor_torchaudio.TorchMFCC": "tssep_b200.feature_extractor_torchaudio.TorchMFCC",
    "tssep.train.net.MaskEstimator_v2": "tssep_b200.net.MaskEstimator_v2",
    "tssep.train.net.InstanceNorm": "tssep_b200.net.InstanceNorm",
    "tssep.train.net.InstanceNorm_v2": "tssep_b200.net.InstanceNorm_v2",
    "tssep.train.rnnp.RNNP_packed": "tssep_b200.rnnp.RNNP_packed",
    "tssep.train.enhancer.Masking": "tssep_b200.enhancer.Masking",
    "tssep.train.enhancer.TorchBF": "tssep_b200.enhancer.TorchBF",
    "tssep.train.loss.LogMAE": "tssep_b200.loss.LogMAE",
    "tssep.train.loss.MAE": "tssep_b200.loss.MAE",
    "tssep.train.init_ckpt.InitCheckPoint": "tssep_b200.init_ckpt.InitCheckPoint",
    "tssep.train.init_ckpt.InitCheckPointVAD2Sep": "tssep_b200.init_ckpt.InitCheckPointVAD2Sep",
    "tssep.train.loss.VADSigmoidBCE": "tssep_b200.loss.VADSigmoidBCE",
    "tssep.data.DummyReader": "tssep_b200.data.DummyReader",
}


def class_to_str(cls) -> str:
    if isinstance(cls, str):
        return cls
    return f"{cls.__module__}.{cls.__qualname__}"


def import_class(name):
    if not isinstance(name, str):
        return name
    name = FACTORY_ALIASES.get(name, name)  # a reference factory string names this package's drop-in class
    module, _, attr = name.rpartition(".")
    if not module:
        raise ImportError(f"factory {name!r} is not a dotted path")
    return getattr(importlib.import_module(module), attr)


def remap_factories(config):
    """Returns a deep copy of ``config`` with reference factory names replaced by ours."""
    if isinstance(config, collections.abc.Mapping):
        out = {}
        for k, v in config.items():
            if k == "factory":
                v = FACTORY_ALIASES.get(class_to_str(v), v)
            out[k] = remap_factories(v)
        return out
    if isinstance(config, (list, tuple)):
        return type(config)(remap_factories(v) for v in config)
    return copy.deepcopy(config)


def _signature_defaults(cls) -> Dict[str, Any]:
    sig = inspect.signature(cls)
    out = {}
    for p in sig.parameters.values():
        if p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD):
            continue
        out[p.name] = p.default  # inspect._empty marks "required"
    return out


class _Dogmatic(collections.abc.MutableMapping):
    """Config view handed to ``finalize_dogmatic_config``.

    Values supplied by the user (``updates``) win over values assigned inside
    the hook; assigning a dict with a ``factory`` key expands it immediately so
    the hook can instantiate it (``Configurable.from_config(config['fe'])``,
    tssep/train/model.py:137).
    """

    def __init__(self, data, updates):
        self._data = data
        self._updates = updates

    def _merge(self, key, value):
        upd = self._updates.get(key, inspect._empty)
        if isinstance(value, collections.abc.Mapping) and "factory" in value:
            sub_updates = {}
            if isinstance(upd, collections.abc.Mapping):
                sub_updates = dict(upd)
            value = dict(value)
            factory = sub_updates.pop("factory", value.pop("factory"))
            merged = {**value, **sub_updates}
            return _get_config(import_class(factory), merged)
        if upd is not inspect._empty and not isinstance(upd, collections.abc.Mapping):
            return upd
        if upd is not inspect._empty and isinstance(upd, collections.abc.Mapping) and "factory" in upd:
            upd = dict(upd)
            return _get_config(import_class(upd.pop("factory")), upd)
        return value

    def __setitem__(self, key, value):
        self._data[key] = self._merge(key, value)

    def __getitem__(self, key):
        return self._data[key]

    def __delitem__(self, key):
        del self._data[key]

    def __iter__(self):
        return iter(self._data)

    def __len__(self):
        return len(self._data)


def _get_config(cls, updates=None) -> Dict[str, Any]:
    updates = dict(updates or {})
    updates.pop("factory", None)
    defaults = _signature_defaults(cls)
    unknown = set(updates) - set(defaults)
    if unknown:
        raise TypeError(f"{class_to_str(cls)} got unexpected config keys {sorted(unknown)}")
    data = {}
    view = _Dogmatic(data, updates)
    for k, v in defaults.items():
        if k in updates:
            view[k] = updates[k] if not (isinstance(updates[k], collections.abc.Mapping) and "factory" not in updates[k]) else v
        else:
            data[k] = v
    hook = getattr(cls, "finalize_dogmatic_config", None)
    if hook is not None:
        hook(view)
    # a user update that is a plain dict (no factory) refines a factory dict set by the hook
    for k, upd in updates.items():
        if isinstance(upd, collections.abc.Mapping) and "factory" not in upd:
            cur = data.get(k)
            if isinstance(cur, collections.abc.Mapping) and "factory" in cur:
                merged = {**{kk: vv for kk, vv in cur.items() if kk != "factory"}, **upd}
                data[k] = _get_config(import_class(cur["factory"]), merged)
            else:
                data[k] = dict(upd)
    missing = [k for k, v in data.items() if v is inspect._empty]
    if missing:
        raise TypeError(f"{class_to_str(cls)}: missing config values for {missing}")
    return {"factory": class_to_str(cls), **{k: data[k] for k in defaults}}


def _instantiate(config):
    if isinstance(config, collections.abc.Mapping):
        if "factory" in config:
            cls = import_class(config["factory"])
            kwargs = {k: _instantiate(v) for k, v in config.items() if k != "factory"}
            return cls(**kwargs)
        return {k: _instantiate(v) for k, v in config.items()}
    if isinstance(config, (list, tuple)):
        return type(config)(_instantiate(v) for v in config)
    return config


class Configurable:
    """Mix-in giving a class the reference's config protocol."""

    @classmethod
    def finalize_dogmatic_config(cls, config):
        pass

    @classmethod
    def get_config(cls, updates=None):
        return _get_config(cls, updates)

    @classmethod
    def from_config(cls, config):
        obj = _instantiate(config)
        return obj

    @classmethod
    def new(cls, updates=None):
        return cls.from_config(cls.get_config(updates))
