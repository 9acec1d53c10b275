"""Checkpoint / weight interop of the drop-in model.

* ``load_model_weights``   raw ``state_dict`` files (``*_best.pth``) or full training checkpoints that wrap it under
                           ``"model_state"`` (reference: src/scripts/evaluate.py:259-267).
* ``load_pretrained``      warm start: only tensors whose name AND shape match are taken, so a checkpoint trained with a
                           different ``n_classes`` transfers everything but the two class heads
                           (reference: src/scripts/train.py:126-176).
* ``save_packed`` / ``load_packed``  the engine's one-time fold + pack (BN folded in fp32, bf16 K-major GEMM weights,
                           fused-block side tables, merged q|k|v, stem GEMM weights) cached next to the checkpoint, keyed
                           by a digest of the ``state_dict`` it was made from, the precision and the C-ABI version, so
                           that N evaluation processes (one per GPU) do not each redo it.  A stale or foreign cache is
                           ignored, never trusted: the file holds tensors and plain containers only and is read with
                           ``weights_only=True`` (no unpickling of arbitrary objects), and the header is compared before use.
"""

from __future__ import annotations

from pathlib import Path
from typing import Any, Dict, List, Tuple

import torch

from .synthetic import state_dict_digest

PACK_FORMAT = 3


def load_model_weights(checkpoint_path, device="cpu") -> Dict[str, Any]:
    """-> the model ``state_dict`` of a raw or wrapped checkpoint file (reference: evaluate.py:259-267)."""
    ckpt = torch.load(Path(checkpoint_path), map_location=device, weights_only=True)
    if isinstance(ckpt, dict) and "model_state" in ckpt:
        return ckpt["model_state"]
    return ckpt


def load_pretrained(model: torch.nn.Module, checkpoint_path, device="cpu") -> Tuple[List[str], List[str], List[str]]:
    """Warm-start ``model`` from a checkpoint of a possibly different dataset (reference: train.py:126-176).

    -> (loaded, skipped_shape_mismatch, skipped_unknown) key lists."""
    pretrained = load_model_weights(checkpoint_path, device)
    state = model.state_dict()
    loaded = [k for k, v in pretrained.items() if k in state and v.shape == state[k].shape]
    mismatch = [k for k in pretrained if k in state and pretrained[k].shape != state[k].shape]
    unknown = [k for k in pretrained if k not in state]
    state.update({k: pretrained[k] for k in loaded})
    model.load_state_dict(state)
    return loaded, mismatch, unknown


def packed_cache_path(checkpoint_path, precision: str = "bf16") -> Path:
    p = Path(checkpoint_path)
    return p.with_name(p.name + f".cabinet_b200.{precision}.pack")


def _header(model) -> Dict[str, Any]:
    from . import _lib

    return {"format": PACK_FORMAT, "abi": int(_lib.load().cabinet_abi_version()), "precision": model.precision,
            "n_classes": model.n_classes, "mode": model.mode, "digest": state_dict_digest(model.state_dict())}


def save_packed(model, path) -> Path:
    """Write the packed weights of ``model``'s current engine (CUDA model) to ``path``."""
    eng = model.engine()
    blob = {"header": _header(model), "packed": eng.packed_state()}
    path = Path(path)
    tmp = path.with_name(path.name + ".tmp")
    torch.save(blob, tmp)
    tmp.replace(path)  # atomic: concurrent ranks either see a complete file or none
    return path


def load_packed(model, path) -> bool:
    """Install cached packed weights into ``model`` (on its CUDA device) if ``path`` holds a pack made from exactly
    these weights / precision / ABI.  -> True when used; False (and the engine packs lazily as usual) otherwise."""
    from .engine import Engine

    path = Path(path)
    if not path.is_file():
        return False
    dev = next(model.parameters()).device
    try:
        blob = torch.load(path, map_location=dev, weights_only=True)
    except Exception:
        return False
    if not isinstance(blob, dict) or blob.get("header") != _header(model):
        return False
    eng = Engine(model, precision=model.precision, packed=blob["packed"])
    eng.stamp = (model.precision, model._weights_stamp())
    model.__dict__["_engine"] = eng
    return True
