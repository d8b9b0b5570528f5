"""accflow_b200 — B200-native flow-estimation + backward-accumulation path of AccFlow.

Public surface mirrors the reference's ``networks`` package:
``accflow_b200.networks.build_flow_estimator`` and ``accflow_b200.networks.AccFlow_.AccFlow``.
"""
__version__ = "0.1.0"
