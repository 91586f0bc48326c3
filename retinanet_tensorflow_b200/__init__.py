"""retinanet_tensorflow_b200 -- the RetinaNet dense-head data path of vshmyhlo/retinanet-tensorflow
(anchors -> assignment/encoding -> losses fwd+bwd -> decode/NMS) as hand-written sm_100a CUDA
kernels behind a C ABI (include/rndense.h), with a thin Python host that keeps the reference's
module surface: `levels`, `dataset`, `utils`, `losses` (+ `pipeline.DenseHeadPath`, the fused
batched form).  There is no CPU fallback: importing the compute modules without the built
library raises."""

__all__ = ['levels', 'dataset', 'utils', 'losses', 'augmentation', 'metrics', 'pipeline', 'dist', 'synthetic']
