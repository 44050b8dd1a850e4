"""Drop-in for the inference-path names of the reference's train_autoencoder.py."""
import cic_b200 as _cic
from cic_b200.autoencoder import build_autoencoder, load_images_from_folder  # noqa: F401


def main():
    raise NotImplementedError("autoencoder training (train_autoencoder.py:58-90) is outside the accelerated "
                              "inference path (SURVEY.md §2 #20)")


if __name__ == "__main__":
    main()
