"""psnerf_b200 — B200-native (sm_100a) implementation of the PS-NeRF render / shading hot path.

Drop-in host modules (same constructors, forward signatures and state-dict keys as ywq/psnerf):
    psnerf_b200.stage1.NeuralNetwork / Renderer      <- stage1/model/{network,rendering}.py
    psnerf_b200.stage2.PSNetwork                     <- stage2/model/renderer.py
All arithmetic runs in the hand-written CUDA library behind include/psnerf_b200.h; there is no CPU path.
"""
__version__ = "0.1.0"
