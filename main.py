"""Entry point with the reference's CLI: ``python main.py --output_dir runs/ --epochs 200``
(multi-GPU: ``torchrun --nproc-per-node N main.py ...``)."""
from tf2_cyclegan_b200.cli import main

if __name__ == "__main__":
  main()
