int oracle_placeholder(void){return 0;}
