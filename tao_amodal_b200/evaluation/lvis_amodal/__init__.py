import logging
from .lvis import LVIS
from .results import LVISResults
from .eval import LVISEval

# same side effect as the reference package (tao_amodal/evaluation/lvis_amodal/__init__.py:7-10);
# LVISVis (matplotlib drawing, lvis_amodal/vis.py) is out of scope and not re-exported
logging.basicConfig(
    format="[%(asctime)s] %(name)s %(levelname)s: %(message)s", datefmt="%m/%d %H:%M:%S",
    level=logging.WARN,
)

__all__ = ["LVIS", "LVISResults", "LVISEval"]
