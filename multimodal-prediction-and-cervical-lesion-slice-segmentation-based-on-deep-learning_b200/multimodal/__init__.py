"""B200 path of the multimodal severity classifier."""
