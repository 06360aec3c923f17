#!/usr/bin/env python
"""TEST INFRASTRUCTURE (development container only): make the reference emit the generated C for the UR5
functions the CPU baseline loops over (J/Tx of EE, M, g, C, R of EE) into $HOME/.cache/abr_control/ur5/."""
import numpy as np
from abr_control.arms import ur5

rc = ur5.Config()
q = np.full(6, 0.1)
rc.J("EE", q), rc.Tx("EE", q), rc.M(q), rc.g(q), rc.C(q, q), rc.R("EE", q)
print(rc.config_folder)
