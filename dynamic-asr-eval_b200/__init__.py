"""dae — B200-native kernels behind the dynamic-evaluation hot path of dynamic-asr-eval.

Python call sites keep the reference's signatures:
    dae.CTCLoss, dae.SpecAugment, dae.GreedyCTCDecoder, dae.SoftDTW, dae.BeamSearch,
    dae.word_error_rate_detail, and dae.lib.{dynamic_eval, AWMC, prepare_chunks, ...}.
Everything numeric runs in libdae.so (hand-written sm_100a CUDA behind a C ABI).
"""
__version__ = "0.1.0"
