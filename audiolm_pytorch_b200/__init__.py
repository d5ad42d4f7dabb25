"""B200-native AudioLM hot path (SoundStream codec convs + RVQ, Semantic/Coarse/Fine transformers).

Same class names, constructor kwargs, forward()/generate()/tokenize() signatures and state_dict keys
as lucidrains/audiolm-pytorch; the arithmetic underneath is hand-written sm_100a CUDA reached
through the C ABI in include/alm_b200.h (libalm_b200.so).
"""
__version__ = "0.1.0"
