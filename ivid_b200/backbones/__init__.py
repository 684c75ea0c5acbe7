from .adm import AdmUnet2d
