"""Black-box model wrapper behind the reference's ``Model`` interface (``model.py:11-328``)."""
import torch

from .lowering import lower


class Model:
    def __init__(self, arch):
        self.arch = arch

    def get_hops(self, num_relations=0):
        """model.py:28-60: number of MessagePassing modules, integer-divided by #relations."""
        num_hops = lower(self.arch).n_message_passing
        if num_relations > 0:
            num_hops //= num_relations
        return num_hops

    def infer(self, *a, **k):
        raise NotImplementedError("inference on a materialised batch (model.py:62-116) is replaced by "
                                  "engine.MaskedForward; there is no CPU path")

    predict_hetero_output = infer

    @staticmethod
    def hetero2homo_output(hetero_output):
        """model.py:255-292."""
        if isinstance(hetero_output, torch.Tensor):
            return hetero_output, None
        values = list(hetero_output.values())
        types = torch.hstack([torch.zeros(len(v), device=v.device, dtype=torch.int) + i for i, v in enumerate(values)])
        return torch.vstack(values), types

    @staticmethod
    def extract_node_edge_output(output, ind, n):
        """model.py:294-328."""
        return output[torch.arange(start=ind, end=output.shape[0], step=n)]
