"""Encoder / predictor modules that feed the joint (reference networks/encoder.py:20-108 and
networks/decoder.py:21-126).  The RNN stacks themselves are library cuDNN calls and out of scope
(SURVEY.md 2); these thin modules exist so that ``JointNet(transnet_params, prednet_params,
num_classes)`` keeps the reference constructor and **state_dict keys** (``encoder.rnn.*``,
``encoder.out_proj.*``, ``decoder.embedding.*``, ``decoder.rnn.*``, ``decoder.out_proj.*``), and to
fix the length plumbing of rows A6/A9: lengths arrive as the host lists the collate already makes
(dataloader.py:20-24,37) and go straight into ``pack_padded_sequence(enforce_sorted=False)`` -- no
CPU sort, no index H2D copy, no double gather per step (encoder.py:93-102, decoder.py:103-120).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn
from torch import Tensor
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

_RNNS = {"lstm": nn.LSTM, "gru": nn.GRU, "rnn": nn.RNN}
Lengths = Union[Sequence[int], Tensor, None]


def _host_lengths(lengths: Lengths) -> Optional[Tensor]:
    """int64 CPU tensor from a host list (no device traffic).  A CUDA tensor would force a D2H
    sync, which is exactly the round-trip the reference README.md:65 complains about -> rejected."""
    if lengths is None:
        return None
    if isinstance(lengths, Tensor):
        if lengths.device.type != "cpu":
            raise RuntimeError("sequence lengths for pack_padded_sequence must live on the host "
                               "(pass the collate's Python list); device lengths go to RNNTLoss")
        return lengths.to(torch.int64)
    return torch.as_tensor(list(lengths), dtype=torch.int64)


def _run_packed(rnn, inputs: Tensor, lengths: Optional[Tensor], state=None):
    if lengths is None:
        return rnn(inputs, state)
    packed = pack_padded_sequence(inputs, lengths, batch_first=True, enforce_sorted=False)
    out, state = rnn(packed, state)
    # like the reference (encoder.py:101): padded back to the longest sequence of the batch
    out, _ = pad_packed_sequence(out, batch_first=True)
    return out, state


class AudioTransNet(nn.Module):
    """Transcription network: (bi)RNN stack + ``out_proj`` (reference encoder.py:45-76)."""

    supported_rnns = _RNNS

    def __init__(self, input_size: int, hidden_size: int, output_size: int, num_layers: int,
                 rnn_type: str = "lstm", dropout: float = 0.2, bidirectional: bool = True):
        super().__init__()
        self.hidden_size = hidden_size
        self.rnn = _RNNS[rnn_type.lower()](
            input_size=input_size, hidden_size=hidden_size, num_layers=num_layers, bias=True,
            batch_first=True, dropout=(dropout if num_layers > 1 else 0.0),
            bidirectional=bidirectional)
        self.out_proj = nn.Linear(2 * hidden_size if bidirectional else hidden_size, output_size)

    def forward(self, inputs: Tensor, inputs_lengths: Lengths) -> Tensor:
        out, _ = _run_packed(self.rnn, inputs, _host_lengths(inputs_lengths))
        return self.out_proj(out)


class TextPredNet(nn.Module):
    """Prediction network: Embedding(padding_idx=blank) + RNN + ``out_proj`` (decoder.py:57-80)."""

    supported_rnns = _RNNS

    def __init__(self, embedding_size: int, pad_token_id: int, hidden_size: int, output_size: int,
                 num_layers: int, rnn_type: str = "lstm", dropout: float = 0.2):
        super().__init__()
        self.hidden_size = hidden_size
        self.embedding = nn.Embedding(embedding_size, hidden_size, padding_idx=pad_token_id)
        self.rnn = _RNNS[rnn_type.lower()](
            input_size=hidden_size, hidden_size=hidden_size, num_layers=num_layers, bias=True,
            batch_first=True, dropout=(dropout if num_layers > 1 else 0.0), bidirectional=False)
        self.out_proj = nn.Linear(hidden_size, output_size)

    def forward(self, inputs: Tensor, input_lengths: Lengths = None,
                prev_hidden_state=None) -> Tuple[Tensor, Tensor]:
        embedded = self.embedding(inputs)
        out, hidden = _run_packed(self.rnn, embedded, _host_lengths(input_lengths), prev_hidden_state)
        return self.out_proj(out), hidden
