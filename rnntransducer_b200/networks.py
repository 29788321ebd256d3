"""Encoder / predictor modules that feed the joint (reference networks/encoder.py:20-108 and
networks/decoder.py:21-126).  The RNN stacks themselves are library cuDNN calls and out of scope
(SURVEY.md 2); these thin modules exist so that ``JointNet(transnet_params, prednet_params,
num_classes)`` keeps the reference constructor and **state_dict keys** (``encoder.rnn.*``,
``encoder.out_proj.*``, ``decoder.embedding.*``, ``decoder.rnn.*``, ``decoder.out_proj.*``), and to
restructure the packed-RNN call (rows A6/A9).

What the reference does per network and step (encoder.py:93-102, decoder.py:103-120): lengths list ->
CPU tensor -> ``torch.sort`` -> gather the batch -> ``pack_padded_sequence`` -> RNN ->
``pad_packed_sequence`` -> ``torch.sort`` of the permutation -> gather back.  What is needed instead
depends only on facts the HOST already has (the collate's Python lists, dataloader.py:20-24,37):

* **unidirectional RNN** (the predictor, always; the encoder with ``bidirectional=False``): an output at
  step t depends on steps <= t only, so padding after a sequence's end cannot reach its valid outputs:
  no sort, no gather, no pack -- the RNN runs on the padded batch and the padded outputs are zeroed with a
  device-side mask (what ``pad_packed_sequence`` would have produced, so ``out_proj`` sees the same input).
* **all sequences equally long**: nothing to pack for any RNN.
* **bidirectional RNN, lengths already descending** (what ``data.collate_sorted`` /
  ``data.DistributedBucketSampler`` deliver): ``pack_padded_sequence(enforce_sorted=True)`` -- no sort, no
  index upload, no gather in either direction.
* otherwise: ``enforce_sorted=False`` (torch sorts on the host and index-selects on the device).

Results are those of the reference on the same inputs in every case (tests/test_host.py pins all four
against the reference's own modules).  One documented difference: on the unpacked unidirectional path the
returned final hidden state is that of the padded run (exact only for full-length rows); the training
step discards it (transducer.py:90) and the decode loop passes no lengths.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn
from torch import Tensor
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

_RNNS = {"lstm": nn.LSTM, "gru": nn.GRU, "rnn": nn.RNN}
Lengths = Union[Sequence[int], Tensor, None]


def _host_lengths(lengths: Lengths) -> Optional[List[int]]:
    """Python ints from the collate's list.  A CUDA tensor would force a D2H sync, which is exactly
    the round-trip the reference README.md:65 complains about -> rejected."""
    if lengths is None:
        return None
    if isinstance(lengths, Tensor):
        if lengths.device.type != "cpu":
            raise RuntimeError("sequence lengths for the RNN stacks must live on the host "
                               "(pass the collate's Python list); device lengths go to RNNTLoss")
        return [int(v) for v in lengths.tolist()]
    return [int(v) for v in lengths]


def _length_mask(lengths: List[int], max_len: int, device, dtype) -> Tensor:
    """[B, max_len, 1] mask of valid steps, built on the device from one tiny non-blocking upload."""
    lens = torch.tensor(lengths, dtype=torch.int32)
    if device.type == "cuda":
        lens = lens.pin_memory().to(device, non_blocking=True)
    return (torch.arange(max_len, device=device, dtype=torch.int32)[None, :] < lens[:, None]).unsqueeze(-1).to(dtype)


def _run_rnn(rnn, inputs: Tensor, lengths: Optional[List[int]], state=None):
    """See the module docstring.  Output is padded to the longest sequence of the batch, zeros after each
    sequence's end -- exactly ``pad_packed_sequence``'s result (encoder.py:101)."""
    if lengths is None:
        return rnn(inputs, state)
    if len(lengths) != inputs.size(0):
        raise RuntimeError("one length per sequence expected")
    longest, shortest = max(lengths), min(lengths)
    if shortest <= 0 or longest > inputs.size(1):
        raise RuntimeError("sequence lengths must be in [1, padded length]")
    if longest < inputs.size(1):
        inputs = inputs[:, :longest]
    if shortest == longest:                                   # nothing to pack
        return rnn(inputs, state)
    if not rnn.bidirectional:                                 # causal: padding cannot reach valid outputs
        out, state = rnn(inputs, state)
        return out * _length_mask(lengths, longest, out.device, out.dtype), state
    is_sorted = all(lengths[i] >= lengths[i + 1] for i in range(len(lengths) - 1))
    packed = pack_padded_sequence(inputs, torch.tensor(lengths, dtype=torch.int64), batch_first=True,
                                  enforce_sorted=is_sorted)
    out, state = rnn(packed, state)
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
        out, _ = _run_rnn(self.rnn, inputs, _host_lengths(inputs_lengths))
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
        out, hidden = _run_rnn(self.rnn, embedded, _host_lengths(input_lengths), prev_hidden_state)
        return self.out_proj(out), hidden
