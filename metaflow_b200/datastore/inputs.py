"""`inputs` object handed to join steps (surface of metaflow/datastore/inputs.py:1-17)."""


class Inputs(object):
    """foreach join: inputs[0].x / for inp in inputs; static split join: inputs.<step_name>.x"""

    def __init__(self, flows):
        self.flows = list(flows)
        for f in self.flows:
            setattr(self, f._current_step, f)

    def __getitem__(self, idx):
        return self.flows[idx]

    def __iter__(self):
        return iter(self.flows)

    def __len__(self):
        return len(self.flows)
