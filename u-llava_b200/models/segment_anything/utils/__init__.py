"""models.segment_anything.utils of the B200 build: only transforms.ResizeLongestSide is on the u-LLaVA path
(dataset/tools/mask_toolbox.py:5,13); amg / onnx helpers of the reference package are out of scope."""
