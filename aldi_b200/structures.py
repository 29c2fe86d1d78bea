"""Boxes / Instances containers with the Detectron2 field names the ALDI layer touches
(aldi/pseudolabeler.py:51-67 builds `Instances(image_size){gt_boxes, gt_classes, scores}` on the CPU)."""
import torch


class Boxes:
    def __init__(self, tensor):
        tensor = torch.as_tensor(tensor, dtype=torch.float32)
        if tensor.numel() == 0:
            tensor = tensor.reshape(-1, 4)
        assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
        self.tensor = tensor

    def to(self, device):
        return Boxes(self.tensor.to(device))

    def __len__(self):
        return self.tensor.shape[0]

    def __getitem__(self, item):
        t = self.tensor[item]
        return Boxes(t.view(1, -1) if t.dim() == 1 else t)

    def area(self):
        b = self.tensor
        return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])


class Instances:
    def __init__(self, image_size, **fields):
        object.__setattr__(self, "_image_size", image_size)
        object.__setattr__(self, "_fields", {})
        for k, v in fields.items():
            self.set(k, v)

    @property
    def image_size(self):
        return self._image_size

    def __setattr__(self, name, val):
        self.set(name, val)

    def __getattr__(self, name):
        f = object.__getattribute__(self, "_fields")
        if name not in f:
            raise AttributeError("Cannot find field '{}' in the given Instances!".format(name))
        return f[name]

    def set(self, name, value):
        if len(self._fields):
            assert len(self) == len(value), "Adding a field of length {} to a Instances of length {}".format(
                len(value), len(self))
        self._fields[name] = value

    def has(self, name):
        return name in self._fields

    def get_fields(self):
        return self._fields

    def to(self, device):
        return Instances(self._image_size, **{k: (v.to(device) if hasattr(v, "to") else v) for k, v in self._fields.items()})

    def __len__(self):
        for v in self._fields.values():
            return len(v)
        raise NotImplementedError("Empty Instances does not support __len__!")
