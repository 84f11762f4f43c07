"""Mirror of components/style_transfer_depth/style_transfer_depth.py: DepthStyle (same methods and return values).

The reference loads Depth-Anything-V2 through transformers.pipeline in the constructor (:27); there is no network here, so
the estimator is a constructor argument: any callable image -> {"depth": PIL 'L' image} (the pipeline's own contract).
Without one, the constructor tries the reference's pipeline call and lets its error through."""
import numpy as np

from .Style_a3 import StyleA3
from .util import generate_mip_layers, reconstruct_mip_image


class DepthStyle:
    def __init__(self, device="cuda", depth_pipeline=None):
        if depth_pipeline is None:
            from transformers import pipeline
            depth_pipeline = pipeline(task="depth-estimation", model="depth-anything/Depth-Anything-V2-Small-hf")   # :27
        self.depth_pipeline = depth_pipeline
        self.style_model = StyleA3(device=device, depth_pipeline=depth_pipeline)                                     # :29
        self.style_pipeline = self.style_model.style_transfer

    def get_depth_map(self, image):
        """:33-44"""
        return np.asarray(self.depth_pipeline(image)["depth"])

    def style_transfer(self, image, style, strength=1):
        """:46-59"""
        return self.style_pipeline(style, image, strength=strength)

    def process_mip_layers(self, masked_images, style):
        """:61-72: plane ind is stylised with strength 1 - ind / n; the planes share the style targets and run at the same time
        (one plan / stream each, nst_run_frames_host)."""
        strengths = [(1 - ind / len(masked_images)) for ind in range(len(masked_images))]
        return self.style_model.style_transfer_planes(style, masked_images, strengths)

    def style_MIP(self, image, style, n=2):
        """:74-90"""
        depth = self.get_depth_map(image)
        masked_images = generate_mip_layers(image, depth, n)
        stylized_images = self.process_mip_layers(masked_images, style)
        final_image = reconstruct_mip_image(stylized_images, depth, n)
        return final_image, stylized_images

    def style_Dept(self, image, style):
        """:92-105"""
        return self.style_pipeline(style, image, depth=True)

    def depth_split(self, image, n=2):
        """:107-119"""
        depth = self.get_depth_map(image)
        return generate_mip_layers(image, np.asarray(depth), n)

    def close(self):
        self.style_model.close()
