# round 2: the torch-free multi-GPU driver at world 2 after the host-step protocol change (verdict collective on every rank and every call)
mkdir -p gpurun_out
g++ -std=c++17 -O1 -Iinclude tests/cpp/dist_driver.cpp -o /tmp/dist_driver -Lmvtopicmodel_b200 -lmvtm -lpthread -Wl,-rpath,$PWD/mvtopicmodel_b200 && timeout 40 /tmp/dist_driver 2 6 2 > gpurun_out/r2_dist_driver_world2_final.log 2>&1; echo "driver rc=$?"; tail -3 gpurun_out/r2_dist_driver_world2_final.log
