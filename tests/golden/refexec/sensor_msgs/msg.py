"""Stub of the ROS message module imported by the reference's utils/util_functions.py:3."""


class JointState:
    pass
