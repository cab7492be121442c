from .model import model_factory  # noqa: F401
from .rp_net import RP_Net, ContextCorrelationEncoder, Correlation, dice_ce, dice_loss_softmax  # noqa: F401
from .unet import U_Net  # noqa: F401
from .vgg import Encoder  # noqa: F401
from .modules import conv_block, up_conv  # noqa: F401
