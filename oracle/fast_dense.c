int kpc_fastdense_placeholder(void){return 0;}
