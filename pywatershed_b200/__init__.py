"""pywatershed_b200: B200-native (sm_100a CUDA, fp64) implementation of
pywatershed's per-timestep PRMS/NHM process chain, behind the reference's
Process / Model API.  See DESIGN.md and INTEGRATION.md."""
from .budget import Budget  # noqa: F401
from .control import Control  # noqa: F401
from .engine import Engine, Registry, calendar  # noqa: F401
from .flow_graph import (  # noqa: F401
    FlowGraph,
    FlowNode,
    FlowNodeMaker,
    HruNodeFlowExchange,
    HruSegmentFlowAdapter,
    ObsInFlowNodeMaker,
    ObsInNodeMaker,
    PassThroughFlowNodeMaker,
    PassThroughNodeMaker,
    PRMSChannelFlowNodeMaker,
    SourceSinkFlowNodeMaker,
    prms_channel_flow_graph_postprocess,
    prms_channel_flow_graph_to_model_dict,
)
from .model import Model  # noqa: F401
from .parameters import Parameters, PrmsParameters  # noqa: F401
from .processes import (  # noqa: F401
    ConservativeProcess,
    PRMSAtmosphere,
    PRMSCanopy,
    PRMSChannel,
    PRMSEt,
    PRMSGroundwater,
    PRMSGroundwaterNoDprst,
    PRMSRunoff,
    PRMSRunoffNoDprst,
    PRMSSnow,
    PRMSSoilzone,
    PRMSSoilzoneNoDprst,
    PRMSSolarGeometry,
    Process,
    TimeseriesArray,
)

__all__ = [
    "Budget", "Control", "Engine", "Registry", "calendar", "Model", "Parameters",
    "PrmsParameters", "Process", "ConservativeProcess", "TimeseriesArray",
    "PRMSSolarGeometry", "PRMSAtmosphere", "PRMSCanopy", "PRMSSnow", "PRMSRunoff",
    "PRMSSoilzone", "PRMSGroundwater", "PRMSChannel", "PRMSRunoffNoDprst", "PRMSSoilzoneNoDprst",
    "PRMSGroundwaterNoDprst", "PRMSEt", "FlowGraph", "FlowNode", "FlowNodeMaker",
    "PRMSChannelFlowNodeMaker", "PassThroughFlowNodeMaker", "PassThroughNodeMaker",
    "ObsInFlowNodeMaker", "ObsInNodeMaker", "SourceSinkFlowNodeMaker", "HruSegmentFlowAdapter",
    "prms_channel_flow_graph_postprocess", "prms_channel_flow_graph_to_model_dict", "HruNodeFlowExchange",
]
