"""vggsfm_b200 -- B200-native geometry hot path for VGGSfM."""
