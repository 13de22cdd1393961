from .recognition import RecognitionTask
from .oscc import OSCCTask
from .lta import LTATask
from .pnr import PNRTask

__all__ = ["RecognitionTask", "OSCCTask", "LTATask", "PNRTask"]
