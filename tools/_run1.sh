AB_TIMEOUT=80 bash tools/ab_scale.sh 4 bal nobal:SFE_FANOUT_SPLIT=2 bal2
