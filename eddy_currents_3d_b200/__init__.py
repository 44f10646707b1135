"""B200-native drop-in for the hot path of JNSresearcher/eddy_currents_3d (EC3D):
sparse assembly, per-timestep RHS/history/motion and the BiCGSTAB-with-restart solve."""
from .problem import Problem, Source, plate, MU0_LITERAL  # noqa: F401
from .vxc import load_vxc  # noqa: F401
