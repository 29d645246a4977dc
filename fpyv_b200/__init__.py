"""fpyv_b200: B200-native (sm_100a) batched FPV-drone dynamics, drop-in for the `Drone` dynamics path of
omrijsharon/FpyV.  Importing the package never silently degrades: the CUDA library is loaded on first
use and its absence raises."""
from . import config
from ._lib import FpvError

__all__ = ["config", "FpvError", "BatchedDrone", "Drone", "BatchedRacer", "Racer", "Joystick", "Ground",
           "Cylinder", "Target", "Gate", "BatchedCamera", "World", "Autopilot", "PID", "BatchedAcroDrone", "TwoStreamDrones"]


def __getattr__(name):
    if name in ("BatchedDrone", "Drone"):
        from . import drone
        return getattr(drone, name)
    if name in ("BatchedRacer", "Racer"):
        from . import racer
        return getattr(racer, name)
    if name == "Joystick":
        from .sticks import Joystick
        return Joystick
    if name in ("Ground", "Cylinder", "Target", "Gate", "Trail"):
        from . import objects
        return getattr(objects, name)
    if name in ("BatchedCamera", "World"):
        from . import camera
        return getattr(camera, name)
    if name in ("Autopilot", "PID"):
        from . import autopilot
        return getattr(autopilot, name)
    if name == "TwoStreamDrones":
        from .closed_loop import TwoStreamDrones
        return TwoStreamDrones
    if name == "BatchedAcroDrone":
        from .acro import BatchedAcroDrone
        return BatchedAcroDrone
    raise AttributeError(name)
