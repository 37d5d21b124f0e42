python tools/ab_config.py 2 2 12 main q20202020 q19212218 q18212318 q19222217
python tools/ab_config.py 4 2 12 main q20202020 q19212218 q18212318 q19222217
for ch in 4 6 8; do
  echo "u5 packed ch=$ch chains:"; python tools/run_config.py $ch 44100 48000 60 16 3 | tail -1
  echo "u5 packed ch=$ch imadhi:"; CRB200_LIB=$PWD/variants/u5imadhi/libclownresampler_b200.so python tools/run_config.py $ch 44100 48000 60 16 3 | tail -1
done
for ch in 3 4 5 6 12; do
  echo "general ch=$ch chains:"; python tools/run_config.py $ch 96000 44100 60 8 3 | tail -1
  echo "general ch=$ch imadhi:"; CRB200_NO_CHAINS=1 python tools/run_config.py $ch 96000 44100 60 8 3 | tail -1
done
