"""
graphinvent_b200 -- B200-native (sm_100a) implementation of GraphINVENT's MPNN hot path.

Public surface (mirrors the reference's `gnn` package):
    graphinvent_b200.gnn.mpnn.{GGNN, MNN, AttentionGGNN, EMN}(constants)   drop-in nn.Modules
    graphinvent_b200.dropin.install()       makes `import gnn.mpnn` resolve to the classes above
    graphinvent_b200.functional.kl_loss     fused Workflow.loss (KLDiv on log_softmax)
    graphinvent_b200.parallel               one-allreduce data-parallel training step
"""
__version__ = "0.1.0"
