"""Minimal Detectron2-compatible containers for callers that do not have Detectron2.

The product accepts any duck type with ``image_size``, ``pred_boxes.tensor``, ``scores``,
``pred_classes`` and ``pred_masks`` (real ``detectron2.structures.Instances`` included);
these two classes exist so that the same call surface is available stand-alone.  The
reference reads fields both as attributes and through ``_fields[...]``
(nn_inference.py:326-327, :357, :375-376), so both are provided.
"""
from __future__ import annotations

from typing import Any, Dict, Tuple

import torch


class Boxes:
    """N x 4 float32 XYXY."""

    def __init__(self, tensor):
        t = torch.as_tensor(tensor, dtype=torch.float32)
        if t.numel() == 0:
            t = t.reshape(0, 4)
        if t.dim() != 2 or t.shape[1] != 4:
            raise ValueError(f"Boxes expects N x 4, got {tuple(t.shape)}")
        self.tensor = t

    def __len__(self) -> int:
        return self.tensor.shape[0]

    def __getitem__(self, item) -> "Boxes":
        t = self.tensor[item]
        return Boxes(t.reshape(1, 4) if t.dim() == 1 else t)

    def to(self, *a, **k) -> "Boxes":
        return Boxes(self.tensor.to(*a, **k))

    def clone(self) -> "Boxes":
        return Boxes(self.tensor.clone())

    @property
    def device(self):
        return self.tensor.device


class Instances:
    def __init__(self, image_size: Tuple[int, int], **fields: Any):
        self.__dict__["_image_size"] = (int(image_size[0]), int(image_size[1]))
        self.__dict__["_fields"] = {}
        for k, v in fields.items():
            self.set(k, v)

    @property
    def image_size(self) -> Tuple[int, int]:
        return self._image_size

    def set(self, name: str, value: Any) -> None:
        if self._fields and len(value) != len(self):
            raise ValueError(f"field {name!r} has length {len(value)}, expected {len(self)}")
        self._fields[name] = value

    def get(self, name: str) -> Any:
        return self._fields[name]

    def has(self, name: str) -> bool:
        return name in self._fields

    def remove(self, name: str) -> None:
        del self._fields[name]

    def get_fields(self) -> Dict[str, Any]:
        return self._fields

    def __setattr__(self, name: str, value: Any) -> None:
        if name.startswith("_"):
            self.__dict__[name] = value
        else:
            self.set(name, value)

    def __getattr__(self, name: str) -> Any:
        f = self.__dict__.get("_fields", {})
        if name in f:
            return f[name]
        raise AttributeError(f"Instances has no field {name!r}")

    def __len__(self) -> int:
        for v in self._fields.values():
            return len(v)
        return 0

    def __getitem__(self, item) -> "Instances":
        if isinstance(item, int):
            n = len(self)
            if not -n <= item < n:
                raise IndexError("Instances index out of range")
            item = slice(item % n, item % n + 1)
        return Instances(self._image_size, **{k: v[item] for k, v in self._fields.items()})

    def to(self, *a, **k) -> "Instances":
        return Instances(self._image_size,
                         **{n: (v.to(*a, **k) if hasattr(v, "to") else v)
                            for n, v in self._fields.items()})
