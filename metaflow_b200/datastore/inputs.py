"""`inputs` object handed to join steps (surface of metaflow/datastore/inputs.py:1-17)."""
from collections.abc import Sequence


class Inputs(Sequence):
    """foreach join: `inputs[0].x`, `for inp in inputs`, `len(inputs)`; static split join: `inputs.<step_name>.x`
    (resolved on access from the flows' step names)."""

    def __init__(self, flows):
        self.flows = tuple(flows)

    def __getitem__(self, idx):
        return self.flows[idx]

    def __len__(self):
        return len(self.flows)

    def __getattr__(self, step_name):
        for flow in self.__dict__.get("flows", ()):
            if flow._current_step == step_name:
                return flow
        raise AttributeError("no input from a step named '%s'" % step_name)
