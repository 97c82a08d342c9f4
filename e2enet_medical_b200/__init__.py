"""e2enet_medical_b200 -- B200-native (sm_100a) implementation of E2ENet's per-patch hot path.

Sub-packages mirror the reference's module paths for the three hot-path files:
    network_architecture.unetpp_d        <- e2enet/network_architecture/unetpp_d.py
    network_architecture.neural_network  <- e2enet/network_architecture/neural_network.py
    sparselearning.core_channel          <- e2enet/training/network_training/sparselearning/core_channel.py
Everything computes through libe2enet_b200.so (C ABI in include/e2enet_b200.h); there is no
CPU or eager-PyTorch fallback.
"""
__version__ = "0.1.0"
