from zotmer_b200.cli import main

main()
