"""The oracle's MT19937 restatement vs torch's own CPU generator (the library the reference draws from)."""
import numpy as np
import torch
from torch.utils.data import DataLoader

from oracle.mt19937 import MT19937


def test_seeding_and_state_blob_layout():
    for seed in (0, 2, 12345, 2 ** 32 + 7):
        torch.manual_seed(seed)
        a = MT19937.from_torch_state(torch.get_rng_state().numpy())
        b = MT19937(seed)
        assert np.array_equal(a.state, b.state) and a.pos == b.pos == 624


def test_draw_rules_match_torch():
    for seed in (3, 99):
        torch.manual_seed(seed)
        mt = MT19937(seed)
        assert np.array_equal(torch.randint(0, 2, (700, 13), dtype=torch.bool).numpy(), mt.randint_bool(700, 13))
        for n in (1, 2, 5, 1002, 5000):
            assert np.array_equal(torch.randperm(n).numpy(), mt.randperm(n))
        for n in (1, 15, 16, 17, 4097):  # masks.py draws, then wlm.py:45 init
            w = torch.nn.Linear(n, 1, bias=False).weight.detach().numpy().ravel()
            assert np.array_equal(w, mt.linear_init(n))
        for _ in DataLoader(torch.zeros(6, 2), batch_size=2, num_workers=0):  # wlm.py:210
            pass
        mt.dataloader_iter()
        blob = torch.get_rng_state().numpy()
        assert np.array_equal(mt.to_torch_state(blob), blob)
        assert np.array_equal(torch.randint(0, 2, (64,), dtype=torch.bool).numpy(), mt.randint_bool(1, 64).ravel())
