"""python -m moonsuperresolution_b200 <flags of process_full_tiles.py>"""
from .engine import main

main()
