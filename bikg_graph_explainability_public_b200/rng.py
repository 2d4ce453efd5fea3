"""Device replay of torch's CPU generator (MT19937) -- the stream the reference draws from.

``set_seed`` (reference ``explainer.py:14-22``) seeds the CPU generator with ``seed + 2``; every
mask bit, ``torch.randperm`` and the surrogate's ``nn.Linear`` init then consume it in order
(SURVEY.md section 7 "RNG stream").  ``DeviceStream`` lifts the generator state to the GPU, lets
the mask kernels consume draws there, and hands the advanced state back to torch so that the
host-side ``nn.Linear`` init continues the very same stream.
"""
import struct

import numpy as np
import torch

from . import _lib

_N = 624


class DeviceStream:
    def __init__(self, device):
        self.lib = _lib.load()
        self.device = device
        self._blob = torch.get_rng_state()  # template for the hand-back
        b = self._blob.numpy().tobytes()
        self.seed, left, _seeded, nxt = struct.unpack("<QiiQ", b[:24])
        words = np.frombuffer(b[24:24 + _N * 8], dtype=np.uint64).astype(np.uint32)
        pos = _N if left == 1 else int(nxt)
        self.state = torch.from_numpy(words.view(np.int32).copy()).to(device)
        self.pos = torch.tensor([pos], dtype=torch.int32, device=device)

    def snapshot(self):
        return self.state.clone(), self.pos.clone()

    def restore(self, snap):
        self.state.copy_(snap[0])
        self.pos.copy_(snap[1])

    def draw(self, n):
        """Next ``n`` tempered 32-bit outputs as an int32 device tensor (bit pattern of the u32)."""
        out = torch.empty(max(int(n), 1), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.xpgnn_mt19937_draw(self.state.data_ptr(), self.pos.data_ptr(), out.data_ptr(), int(n),
                                               _lib.stream_ptr()))
        return out

    def replay(self, snap, n):
        """The ``n`` draws that followed ``snap`` (a ``snapshot()``), without disturbing the live stream."""
        keep = self.snapshot()
        self.restore(snap)
        out = self.draw(n)
        self.restore(keep)
        return out

    def hand_back(self):
        """Write the advanced state into torch's CPU generator (``torch.set_rng_state``)."""
        words = self.state.cpu().numpy().view(np.uint32).astype(np.uint64)
        pos = int(self.pos.item())
        left = 1 if pos == _N else _N + 1 - pos
        b = bytearray(self._blob.numpy().tobytes())
        b[:24] = struct.pack("<QiiQ", self.seed, left, 1, pos)
        b[24:24 + _N * 8] = words.tobytes()
        torch.set_rng_state(torch.frombuffer(bytes(b), dtype=torch.uint8).clone())
