from .vggf import (ConvSpec, FCSpec, VGGOracle, VGGSpec, build_oracle, get_spec, vgg16_spec,
                   vgg_funnel_model, vggf_spec)
