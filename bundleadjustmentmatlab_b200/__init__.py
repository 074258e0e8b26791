"""B200-native drop-in for the LM bundle-adjustment hot path of VLG's toolbox/bundle."""
