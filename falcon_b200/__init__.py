"""falcon_b200 -- B200-native fc_consensus (FALCON pre-assembly consensus hot path).

Host side mirrors falcon_kit/falcon_kit.py (ctypes binding) and falcon_kit/mains/consensus.py
(CLI); the arithmetic lives in falcon_b200/csrc (CUDA, sm_100a) behind include/falcon_b200.h.
"""
__version__ = "0.1.0"
