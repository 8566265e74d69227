"""The oracle's Keras primitives (oracle/reference_port.py: lstm_cell, gru_sequence, conv 'same', BatchNormalization, MaxPool1D)
against PyTorch's own layers - an independent implementation of the same published recurrences.  TensorFlow cannot be installed
here, so this does not replace a diff against TF; it rules out slips in the restatement itself (gate algebra, where the reset gate
applies, bias handling, padding sides), with the Keras <-> PyTorch conventions taken from both libraries' documentation:
  LSTM: Keras gate blocks i | f | c | o = PyTorch i | f | g | o; Keras has ONE bias, PyTorch b_ih + b_hh
  GRU:  Keras blocks z | r | h, PyTorch r | z | n; reset_after=True (TF2 default) = PyTorch's form n = tanh(W x + b + r * (U h + b'));
        Keras h' = z h + (1 - z) n = PyTorch (1 - z) n + z h
  Conv 'same', stride 1: both pad total = k - 1 with the extra sample on the right for even k"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import reference_port as O

D = torch.float64


def test_lstm_cell_equals_torch_lstmcell():
    torch.manual_seed(0)
    n_in, u, B = 7, 5, 3
    cell = torch.nn.LSTMCell(n_in, u).to(D)
    x, h, c = torch.randn(B, n_in, dtype=D), torch.randn(B, u, dtype=D), torch.randn(B, u, dtype=D)
    with torch.no_grad():
        h_t, c_t = cell(x, (h, c))
        h_o, c_o = O.lstm_cell(x, h, c, cell.weight_ih.T, cell.weight_hh.T, cell.bias_ih + cell.bias_hh)
    assert torch.allclose(h_o, h_t, atol=1e-12) and torch.allclose(c_o, c_t, atol=1e-12)


def test_gru_reset_after_equals_torch_gru():
    torch.manual_seed(1)
    n_in, u, B, T = 6, 4, 2, 9
    gru = torch.nn.GRU(n_in, u, batch_first=True).to(D)
    x = torch.randn(B, T, n_in, dtype=D)

    def keras_blocks(w):      # PyTorch row blocks r | z | n  ->  Keras column blocks z | r | h
        r, z, n = torch.chunk(w, 3, dim=0)
        return torch.cat([z, r, n], dim=0)

    with torch.no_grad():
        y_t, _ = gru(x)
        kernel = keras_blocks(gru.weight_ih_l0).T
        rec = keras_blocks(gru.weight_hh_l0).T
        bias = torch.stack([keras_blocks(gru.bias_ih_l0), keras_blocks(gru.bias_hh_l0)])
        y_o = O.gru_sequence(x, kernel, rec, bias)
    assert torch.allclose(y_o, y_t, atol=1e-12)


def test_conv_same_stride1_equals_torch_same_padding():
    torch.manual_seed(2)
    for k in (1, 2, 3, 4, 5, 8):
        x = torch.randn(2, 11, 3, dtype=D)
        w = torch.randn(k, 3, 4, dtype=D)
        y = O.conv1d_same_nwc(x, w, None, 1)
        y_t = F.conv1d(x.permute(0, 2, 1), w.permute(2, 1, 0), padding="same").permute(0, 2, 1)
        assert torch.allclose(y, y_t, atol=1e-12), k
    x = torch.randn(2, 6, 7, 3, dtype=D)
    w = torch.randn(3, 3, 3, 5, dtype=D)
    y = O.conv2d_same_nhwc(x, w, 1)
    y_t = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding="same").permute(0, 2, 3, 1)
    assert torch.allclose(y, y_t, atol=1e-12)


def test_conv_same_stride2_geometry():
    """TF's rule for stride 2, k = 3: out = ceil(n / 2); an even size gets its one padding sample at the END, an odd size one on
    each side (what GST.py's six stride-2 layers see for 80 mel bins: 80 -> 40 -> 20 -> 10 -> 5 -> 3 -> 2)."""
    w = torch.zeros(3, 3, 1, 1, dtype=D)
    w[0, 0, 0, 0] = 1.0                         # the kernel picks the top-left sample of every window
    for n, first in ((8, 0), (7, -1)):          # index of the window's first sample for output 0: even -> 0 (no pad in front), odd -> -1
        x = torch.arange(1, n * n + 1, dtype=D).reshape(1, n, n, 1)
        y = O.conv2d_same_nhwc(x, w, 2)
        assert y.shape == (1, -(-n // 2), -(-n // 2), 1)
        assert float(y[0, 0, 0, 0]) == (float(x[0, 0, 0, 0]) if first == 0 else 0.0)
        assert float(y[0, 1, 1, 0]) == float(x[0, 2 + first, 2 + first, 0])
    sizes = [80]
    for _ in range(6):
        sizes.append(-(-sizes[-1] // 2))
    assert sizes == [80, 40, 20, 10, 5, 3, 2]


def test_batchnorm_and_maxpool_equal_torch():
    torch.manual_seed(3)
    x = torch.randn(4, 9, 6, dtype=D)
    g, b, m, v = torch.rand(6, dtype=D) + 0.5, torch.randn(6, dtype=D), torch.randn(6, dtype=D), torch.rand(6, dtype=D) + 0.5
    y = O.batchnorm_inference(x, g, b, m, v)
    y_t = F.batch_norm(x.reshape(-1, 6), m, v, g, b, training=False, eps=1e-3).reshape(4, 9, 6)
    assert torch.allclose(y, y_t, atol=1e-12)
    p = O.max_pool1d_same(x, 2, 1)
    nxt = torch.cat([x[:, 1:], torch.full((4, 1, 6), float("-inf"), dtype=D)], dim=1)
    assert torch.equal(p, torch.maximum(x, nxt))
