"""The Python surface SURVEY 8(b) lists as the drop-in boundary resolves by the reference's module paths and names."""
import importlib

SURFACE = {
    'models': ['create_model'],
    'models.networks': ['define_G', 'define_D', 'define_F', 'init_weights', 'weights_init_kaiming'],
    'models.SRRaGAN_model': ['SRRaGANModel'],
    'models.base_model': ['BaseModel'],
    'models.modules.architecture': ['RRDBNet', 'Discriminator_VGG_128', 'VGGFeatureExtractor'],
    'models.modules.block': ['conv_block', 'act', 'pad', 'get_valid_padding', 'sequential', 'ShortcutBlock', 'RRDB', 'ResidualDenseBlock_5C',
                             'upconv_blcok', 'pixelshuffle_block'],
    'models.modules.loss': ['GANLoss', 'GradientPenaltyLoss', 'CreateRangeLoss', 'FilterLoss', 'Latent_channels_desc_2_num_channels'],
    'CEM.CEMnet': ['CEMnet', 'CEM_PyTorch', 'Filter_Layer', 'CEM_downsampler', 'Get_CEM_Conf', 'Adjust_State_Dict_Keys', 'Return_kernel'],
    'CEM.imresize_CEM': ['imresize', 'calc_strides'],
    'Z_optimization': ['Z_optimizer', 'Optimizable_Z', 'ReturnPatchExtractionMat', 'SoftHistogramLoss', 'TV_Loss', 'ArcTanH'],
}
MODEL_MEMBERS = ['feed_data', 'optimize_parameters', 'test', 'Prepare_Input', 'GetLatent', 'Output_Batch', 'get_current_visuals', 'get_current_log',
                 'perform_validation', 'update_learning_rate', 'get_current_learning_rate', 'save', 'load', 'save_log', 'load_log',
                 'display_log_figure', 'save_network', 'load_network', 'process_loaded_state_dict', 'Set_Require_Grad_Status']


def test_reference_surface_is_present():
    for mod, names in SURFACE.items():
        m = importlib.import_module(mod)
        for n in names:
            assert hasattr(m, n), '%s.%s' % (mod, n)
    from models.SRRaGAN_model import SRRaGANModel
    for n in MODEL_MEMBERS:
        assert callable(getattr(SRRaGANModel, n, None)), n
