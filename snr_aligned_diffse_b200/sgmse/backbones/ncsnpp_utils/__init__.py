"""Operator-level mirror of sgmse-bbed/sgmse/backbones/ncsnpp_utils for the pieces with a public call surface
(`op.upfirdn2d`, `up_or_down_sampling.upsample_2d / downsample_2d`).  The network itself runs in the native executor
(`snr_aligned_diffse_b200.engine`), which uses specialised NHWC kernels for the two configurations NCSN++ needs."""
