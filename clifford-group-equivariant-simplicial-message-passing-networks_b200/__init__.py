"""csmpn_b200 — B200-native hot path of CSMPN (shared simplicial message passing).

Sub-packages mirror the reference's import layout for the path:
  algebra.cliffordalgebra / algebra.metric   <- csmpn/algebra/{cliffordalgebra,metric}.py
  models.cegnn_utils                          <- csmpn/models/cegnn_utils.py
  models.{md17,motion,nba,hulls}_cssmpnn      <- csmpn/models/*_cssmpnn.py
  data.modules.{utils,simplicial_data}        <- csmpn/data/modules/*.py (lifting)
Native code: csrc/ (CUDA, sm_100a) behind the C ABI declared in include/csmpn_b200.h.
"""
__version__ = "0.1.0"
