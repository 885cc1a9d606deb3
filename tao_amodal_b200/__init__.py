"""B200-native TAO-Amodal evaluation hot path (Track-AP + visibility-split frame AP).

Host side: Python mirrors of the reference's ``Tao/TaoResults/TaoEval`` and
``LVIS/LVISResults/LVISEval`` (tao_amodal/evaluation in the reference) over
columnar arrays.  Device side: hand-written sm_100a CUDA behind a C ABI
(``include/ta_eval.h``), loaded with ctypes.  There is no CPU fallback: any
compute call raises if the CUDA library is missing.
"""
__version__ = "0.1.0"
