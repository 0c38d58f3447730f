"""qip_b200 -- a B200-native (sm_100a) state-vector backend behind QIP's StateType interface.

    from qip_b200 import B200Backend
    out, classic = run(node, feed=..., backend_constructor=B200Backend.make_state)

Only the hot path of Renmusxd/QIP is implemented here (state init -> gate apply -> measure);
the graph front-end stays the reference's.  See DESIGN.md and INTEGRATION.md.
"""
from .mats import CMat, SwapMat
from .backend import B200Backend, DeviceState
from .graph import CompiledCircuit, compile_circuit, run

make_state = B200Backend.make_state

__all__ = ["B200Backend", "DeviceState", "CMat", "SwapMat", "make_state", "CompiledCircuit", "compile_circuit", "run"]
