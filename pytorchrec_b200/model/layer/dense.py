"""Dense tower blocks.  Same structure and parameter names as the reference
(torchrec/model/layer/Dense.py:4-24, MLP.py:8-23): ``Dense`` = Linear -> ReLU -> Dropout (the
``activation`` argument is accepted and, as in the reference, always resolves to ReLU);
``MLP`` = ``Sequential`` of ``dense_{i}`` blocks under the attribute ``mlp``.  The GEMMs stay on
cuBLAS through ``nn.Linear`` — library GEMMs are out of this path's scope."""
from typing import List

from torch.nn import Dropout, Linear, Module, ReLU, Sequential


class Dense(Module):
    def __init__(self, input_units: int, output_units: int, activation: str, dropout: float):
        super().__init__()
        self.linear = Linear(input_units, output_units)
        self.activation = ReLU()
        self.dropout = Dropout(dropout)

    def forward(self, x):
        return self.dropout(self.activation(self.linear(x)))


class MLP(Module):
    def __init__(self, input_units: int, hidden_units_list: List[int], activation: str, dropout: float):
        super().__init__()
        self.mlp = Sequential()
        units = input_units
        for index, hidden_units in enumerate(hidden_units_list):
            self.mlp.add_module(f"dense_{index}", Dense(units, hidden_units, activation, dropout))
            units = hidden_units

    def forward(self, x):
        return self.mlp(x)
