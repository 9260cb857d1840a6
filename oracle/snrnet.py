"""CPU restatement of the SNR estimator forward pass (oracle; TEST INFRASTRUCTURE ONLY).

Restates `SNRNet.forward` (sgmse-bbed/sgmse/backbones/snrnet.py:47-97) functionally from a
reference-format state dict (`dnn.<layer>.<param>`), and the SNR branch of `ScoreModel.enhance`
(sgmse-bbed/sgmse/model.py:713-721).
"""
import torch
import torch.nn.functional as F

from . import frontend


def _lstm_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of nn.LSTM (gate order i,f,g,o), batch_first input [B,S,I] -> [B,S,H]."""
    B, S, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    outs = [None] * S
    order = range(S - 1, -1, -1) if reverse else range(S)
    for s in order:
        g = F.linear(x[:, s], w_ih, b_ih) + F.linear(h, w_hh, b_hh)
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs[s] = h
    return torch.stack(outs, dim=1)


def snrnet_forward(sd, x, prefix="dnn."):
    """x: [B,2,256,T] float32 with T % 16 == 0  ->  [B,1] in (0,1) = noise/(speech+noise)."""
    p = prefix
    B = x.shape[0]
    clusters = x.shape[3] // 16
    xs = x.permute(0, 3, 1, 2).reshape(-1, 16, 2, 256).permute(0, 2, 3, 1)  # [B*clusters,2,256,16] :52-54
    f = F.conv2d(xs, sd[p + "conv5x5_1.weight"], sd[p + "conv5x5_1.bias"], padding=2)
    f = F.max_pool2d(f, 2)
    f = F.conv2d(f, sd[p + "conv3x3_1.weight"], sd[p + "conv3x3_1.bias"], padding=1)
    f = F.max_pool2d(f, (2, 1))                                              # [.,32,64,8]
    feats = []
    for i, pool in zip((1, 2, 3, 4), (8, 7, 5, 1)):                          # :65-73
        g = F.conv2d(f, sd[p + f"convt_{i}.weight"], sd[p + f"convt_{i}.bias"])
        feats.append(F.max_pool2d(g, (1, pool)))
    f = torch.cat(feats, dim=1).squeeze(3).squeeze(2).reshape(B, clusters, 128)
    fw = _lstm_dir(f, sd[p + "blstm.weight_ih_l0"], sd[p + "blstm.weight_hh_l0"],
                   sd[p + "blstm.bias_ih_l0"], sd[p + "blstm.bias_hh_l0"], False)
    bw = _lstm_dir(f, sd[p + "blstm.weight_ih_l0_reverse"], sd[p + "blstm.weight_hh_l0_reverse"],
                   sd[p + "blstm.bias_ih_l0_reverse"], sd[p + "blstm.bias_hh_l0_reverse"], True)
    o = torch.cat([fw, bw], dim=2)                                           # [B,clusters,256]
    pooled = torch.cat((o.mean(1), o.std(1), o.min(1)[0], o.max(1)[0]), dim=1)  # std is unbiased :84
    return torch.sigmoid(F.linear(pooled, sd[p + "fc.weight"], sd[p + "fc.bias"]))


def snr_features(y_wave):
    """model.py:715-719: STFT of y/max|y| (no spectrogram transform), re/im as channels, pad to x16."""
    y = y_wave / y_wave.abs().max().item()
    Y = torch.view_as_real(frontend.stft(y)).permute(0, 3, 1, 2)
    return frontend.pad_spec(Y, 16)


def estimate_noise_over_clean(sd, y_wave, prefix="dnn."):
    """model.py:720-721: est_gt = n/(s+n) -> est_snr = est_gt/(1-est_gt) = n/s."""
    with torch.no_grad():
        g = snrnet_forward(sd, snr_features(y_wave), prefix)
    return g / (1 - g)
