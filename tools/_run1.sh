bash tools/ab_scale.sh 2 pub1 pub0:SFE_PUBLISH_IN_SOMA=0
