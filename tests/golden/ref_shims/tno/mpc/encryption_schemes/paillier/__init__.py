from tno.mpc.encryption_schemes.paillier import paillier  # noqa: F401
from tno.mpc.encryption_schemes.paillier.paillier import (  # noqa: F401
    Paillier,
    PaillierCiphertext,
    PaillierPublicKey,
    PaillierSecretKey,
)
