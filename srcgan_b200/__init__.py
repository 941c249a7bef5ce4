"""srcgan_b200 - B200-native (sm_100a) implementation of the SRCGAN G+D training hot path.

Public surface:
  srcgan_b200.nn        RDDBNetB, RDDBNetA (documented shim), NLayerDiscriminator
  srcgan_b200.losses    L1Loss, MSELoss, PSNRLoss, SSIM, DSSIMLoss   (drop-in for src/losses.py)
  srcgan_b200.metrics   MSE, PSNR, AE, SSIM                          (drop-in for src/metrics.py)
  srcgan_b200.color     rgb2lab / lab2rgb on the device
  srcgan_b200.trainer   SRCycleGAN step driver mirroring src/train.py:145-340
  srcgan_b200.data      DevicePrefetcher: the host -> device feed of the training loop, one batch ahead on a side stream
  srcgan_b200.dropin/   directory to put on PYTHONPATH so the reference's unmodified scripts
                        (`import model`, `import losses`, `import metrics`) resolve to this package
"""
__version__ = "0.1.0"
