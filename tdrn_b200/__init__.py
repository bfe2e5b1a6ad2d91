"""tdrn_b200 -- B200 (sm_100a) implementation of the TDRN / DualRefineDet inference hot path.

Drop-in surface (same names and signatures as the reference checkout's top-level packages):
    tdrn_b200.model.dualrefinedet_vggbn.build_net      tdrn_b200.layers.functions.Detect
    tdrn_b200.model.dualrefinedet_mobilenet.build_net  tdrn_b200.layers.functions.PriorBox
    tdrn_b200.model.refinedet_vgg.build_net            tdrn_b200.utils.nms_wrapper.nms
    tdrn_b200.model.ssd4scale_vgg.build_net            tdrn_b200.model.networks.ConvOffset2d
All compute goes through libtdrn_b200.so (include/tdrn_b200.h); there is no CPU fallback.
"""
__version__ = '0.1.0'
