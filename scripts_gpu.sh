cd $GRAFT_REPO_ROOT
./tools/dfma_operands.bin
