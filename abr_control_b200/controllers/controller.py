"""Base class of all controllers (reference: /root/reference/abr_control/controllers/controller.py:4-32)."""
import numpy as np


class Controller:
    def __init__(self, robot_config):
        self.robot_config = robot_config
        self.offset_zeros = np.zeros(3)

    def generate(self, q, dq):
        raise NotImplementedError
