"""GPU: the batched greedy decoder (reference transducer.py:95-145, model.py:76) gives the
tokens of the reference's utterance-by-utterance loop.  On the device the encoder half of the logits comes
from the fused path's projection kernel (rnntb200_joint_cg_project) for all B*T frames at once."""
import pytest
import torch

import rnntransducer_b200 as rb
from test_host import reference_greedy_loop

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["concat_gelu", "add_tanh"])
def test_batched_greedy_decode_on_the_device(cuda_lib, mode):
    torch.manual_seed(4)
    ep = dict(input_size=16, hidden_size=24, output_size=128, num_layers=2, rnn_type="gru", dropout=0.0, bidirectional=True)
    dp = dict(embedding_size=29, pad_token_id=0, hidden_size=24, output_size=128, num_layers=2, rnn_type="lstm", dropout=0.0)
    net = rb.JointNet(ep, dp, 29, mode=mode).cuda().eval()
    with torch.no_grad():
        net.fc.weight.mul_(6.0)
    audio, lengths = torch.randn(5, 23, 16, device="cuda"), [23, 9, 17, 23, 12]
    with torch.no_grad():
        want = reference_greedy_loop(net, audio, lengths, 0, 3)
        got = net.recognize_greedy(audio, lengths, 0, 3)
    assert got.is_cuda and got.dtype == torch.long and sum(len(h) for h in want) > 10
    assert got.shape == (5, max(len(h) for h in want))
    for b, h in enumerate(want):
        assert got[b, :len(h)].tolist() == h and (got[b, len(h):] == 0).all()
    # stopping each row at its own length is a documented option, not the reference's behaviour
    short = net.recognize_greedy(audio, lengths, 0, 3, respect_lengths=True)
    assert short.shape[0] == 5 and short.shape[1] <= got.shape[1]
