from .radius_graph import RadiusGraph
from .lta_temp_connectivity import LTATemporalConnectivity

__all__ = ["RadiusGraph", "LTATemporalConnectivity"]
