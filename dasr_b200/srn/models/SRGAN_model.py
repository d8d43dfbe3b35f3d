"""SRGANModel — reference: codes/SRN/models/SRGAN_model.py (the SRRaGAN step without the relativistic terms:
l_g_gan = w * GAN(D(fake), real), l_d_total = GAN(D(ref), real) + GAN(D(fake), fake), :131-145)."""
from .SRRaGAN_model import SRRaGANModel


class SRGANModel(SRRaGANModel):
    relativistic = False
