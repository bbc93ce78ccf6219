"""Stub of the ROS message package imported by the reference's utils/util_functions.py:4."""


class Vector:
    def __init__(self):
        self.x = []


class MPCData:
    pass
