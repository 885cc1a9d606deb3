import logging
from .tao import Tao
from .results import TaoResults
from .eval import TaoEval

# same side effect as the reference package (tao_amodal/evaluation/tao_amodal/__init__.py:6-10)
logging.basicConfig(
    format="[%(asctime)s] %(name)s %(levelname)s: %(message)s",
    datefmt="%m/%d %H:%M:%S",
    level=logging.WARN,
)

__all__ = ["Tao", "TaoResults", "TaoEval"]
