"""Stub of the ROS service module imported by the reference's utils/util_functions.py:5."""


class Trajectory_Request:
    pass
