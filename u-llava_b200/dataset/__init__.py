"""Host-side mirror of the reference's image preprocessing (dataset/processors/clip_processor.py,
dataset/tools/mask_toolbox.py) running on the device."""
